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


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_multi_gpu_context_matches_single_gpu(mm, small_maps, tiny_maps):
    """mm3d_create_multi: estimateMapsTransforms and composeMaps sharded over two devices inside the library (NCCL feature
    exchange, LPT pair plan, all-to-all of points) are bit-identical to one device."""
    maps, truth = small_maps
    p = mm.default_params(descriptor_type="FPFH")
    one = mm.Context(0)
    two = mm.Context(devices=[0, 1])
    assert two.device_count == 2 and one.device_count == 1
    G1 = one.estimate_maps_transforms(maps, p)
    G2 = two.estimate_maps_transforms(maps, p)
    assert np.array_equal(G1.view(np.uint32), G2.view(np.uint32))
    # more devices than maps with keypoints, an empty map in the middle
    odd = [maps[0], np.zeros((0, 4), np.float32), maps[1]]
    assert np.array_equal(one.estimate_maps_transforms(odd, p).view(np.uint32), two.estimate_maps_transforms(odd, p).view(np.uint32))
    T = np.stack([np.linalg.inv(truth[0]) @ t for t in truth]).astype(np.float32)
    for res in (0.05, 0.2, 1e-4):
        c1 = one.compose_maps(maps, T, res)
        c2 = two.compose_maps(maps, T, res)
        assert c1.shape == c2.shape and np.array_equal(c1.view(np.uint32), c2.view(np.uint32)), res
    T[1] = 0
    assert np.array_equal(one.compose_maps(maps, T, 0.05).view(np.uint32), two.compose_maps(maps, T, 0.05).view(np.uint32))
    one.close(); two.close()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_map_merge_tool_two_gpus(tmp_path, tiny_maps):
    """The CLI on two GPUs (MM3D_DEVICES=0,1) prints the same transforms and writes the same output.pcd as on one."""
    subprocess.check_call(["make", "-C", PKG, "tools"], stdout=subprocess.DEVNULL)
    maps, _ = tiny_maps
    outs = []
    for env_extra in ({"MM3D_DEVICE": "0"}, {"MM3D_DEVICES": "0,1"}):
        d = tmp_path / ("run_" + "_".join(env_extra))
        d.mkdir()
        a, b = str(d / "a.pcd"), str(d / "b.pcd")
        write_pcd(a, maps[0]); write_pcd(b, maps[1])
        env = dict(os.environ); env.update(env_extra)
        r = subprocess.run([os.path.join(PKG, "map_merge_tool"), a, b, "--descriptor_type", "FPFH"], cwd=d, capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stdout + r.stderr
        outs.append((r.stdout.split("Estimated transforms:")[1], open(d / "output.pcd", "rb").read()))
    assert outs[0][0] == outs[1][0]
    assert outs[0][1] == outs[1][1]
