"""map_merge_tool (config 1 of BASELINE.json: two overlapping room scans through the CLI on PCD files)."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "map-merge_b200")


def write_pcd(path, pts, mode="binary"):
    n = len(pts)
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n"
           f"WIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA {mode}\n")
    with open(path, "wb") as f:
        f.write(hdr.encode())
        if mode == "binary":
            f.write(np.ascontiguousarray(pts, np.float32).tobytes())
        else:
            for p in pts:
                f.write(f"{p[0]:.9g} {p[1]:.9g} {p[2]:.9g} {p[3]:.9g}\n".encode())


def read_pcd(path):
    raw = open(path, "rb").read()
    k = raw.index(b"DATA binary\n") + len(b"DATA binary\n")
    n = int(re.search(rb"POINTS (\d+)", raw[:k]).group(1))
    return np.frombuffer(raw[k:k + n * 16], np.float32).reshape(n, 4).copy()


def test_map_merge_tool(tmp_path, ctx, mm, tiny_maps):
    subprocess.check_call(["make", "-C", PKG, "tools"], stdout=subprocess.DEVNULL)
    maps, truth = tiny_maps
    a, b = str(tmp_path / "a.pcd"), str(tmp_path / "b.pcd")
    write_pcd(a, maps[0], "binary")
    write_pcd(b, maps[1], "binary")
    r = subprocess.run([os.path.join(PKG, "map_merge_tool"), a, b, "--descriptor_type", "FPFH", "--unknown_flag", "7"], cwd=tmp_path,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "params: " in r.stdout and "descriptor_type: FPFH" in r.stdout and "resolution: 0.1" in r.stdout
    nums = re.findall(r"[-+]?\d*\.?\d+(?:[eE][-+]?\d+)?", r.stdout.split("Estimated transforms:")[1].split("> Compositing")[0])
    T = np.array(nums, np.float64).reshape(-1, 4, 4)
    p = mm.default_params(descriptor_type="FPFH")
    want = ctx.estimate_maps_transforms(maps, p)
    np.testing.assert_allclose(T, want, atol=2e-5, rtol=1e-5)  # printed with 6 significant digits
    out = read_pcd(str(tmp_path / "output.pcd"))
    comp = ctx.compose_maps(maps, want, 0.05)
    assert out.shape == comp.shape and np.array_equal(out.view(np.uint32), comp.view(np.uint32))
    # fewer than two inputs -> -1 like the reference (map_merge_tool.cpp:14-17)
    r = subprocess.run([os.path.join(PKG, "map_merge_tool"), a], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0 and "Need at least 2 input files" in r.stderr
    # bad enum value -> error exit
    r = subprocess.run([os.path.join(PKG, "map_merge_tool"), a, b, "--keypoint_type", "sift"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0
