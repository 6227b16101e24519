"""PCD reader/writer of the CLI front-end (host only): ascii / binary / binary_compressed round trips."""
import os
import struct
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "map-merge_b200")

SRC = r'''
#include <cstdio>
#include "pcd_io.h"
int main(int argc, char** argv) {
  map_merge_3d::PointCloud c;
  if (mm3d_io::loadPCDFile(argv[1], c) < 0) return 2;
  return mm3d_io::savePCDFileBinary(argv[2], c) == 0 ? 0 : 3;
}
'''


def lzf_literal(data: bytes) -> bytes:
    out = b""
    for i in range(0, len(data), 32):
        chunk = data[i:i + 32]
        out += bytes([len(chunk) - 1]) + chunk
    return out


def test_pcd_round_trips(tmp_path):
    exe = str(tmp_path / "pcd_rt")
    (tmp_path / "rt.cpp").write_text(SRC)
    subprocess.check_call(["g++", "-O1", "-std=c++17", f"-I{ROOT}/include", f"-I{PKG}/host", "-o", exe, str(tmp_path / "rt.cpp")])
    rng = np.random.default_rng(0)
    n = 257
    pts = rng.normal(size=(n, 4)).astype(np.float32)
    pts[:, 3] = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32).view(np.float32)
    # avoid NaN bit patterns in the colour slot for the ascii case
    rgba = (pts[:, 3].view(np.uint32) & 0x00FFFFFF) | 0x3F000000
    pts[:, 3] = rgba.view(np.float32)
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n"
           f"WIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA ")

    def back(path):
        out = str(tmp_path / "out.pcd")
        assert subprocess.call([exe, path, out]) == 0
        raw = open(out, "rb").read()
        k = raw.index(b"DATA binary\n") + 12
        return np.frombuffer(raw[k:], np.float32).reshape(-1, 4)

    p = str(tmp_path / "bin.pcd")
    open(p, "wb").write(hdr.encode() + b"binary\n" + pts.tobytes())
    assert np.array_equal(back(p).view(np.uint32), pts.view(np.uint32))
    p = str(tmp_path / "ascii.pcd")
    open(p, "wb").write(hdr.encode() + b"ascii\n" + "".join(f"{a:.9g} {b:.9g} {c:.9g} {d:.9g}\n" for a, b, c, d in pts).encode())
    assert np.array_equal(back(p).view(np.uint32), pts.view(np.uint32))
    # binary_compressed: field-major payload, LZF (literal runs are valid LZF)
    soa = b"".join(pts[:, k].tobytes() for k in range(4))
    comp = lzf_literal(soa)
    p = str(tmp_path / "comp.pcd")
    open(p, "wb").write(hdr.encode() + b"binary_compressed\n" + struct.pack("<II", len(comp), len(soa)) + comp)
    assert np.array_equal(back(p).view(np.uint32), pts.view(np.uint32))
    # extra fields and rgba as uint32 are tolerated
    hdr2 = hdr.replace("FIELDS x y z rgb", "FIELDS x y z intensity rgba").replace("SIZE 4 4 4 4", "SIZE 4 4 4 4 4") \
        .replace("TYPE F F F F", "TYPE F F F F U").replace("COUNT 1 1 1 1", "COUNT 1 1 1 1 1")
    rows = np.zeros((n, 5), np.float32)
    rows[:, :3] = pts[:, :3]; rows[:, 3] = 7.0; rows[:, 4] = pts[:, 3]
    p = str(tmp_path / "extra.pcd")
    open(p, "wb").write(hdr2.encode() + b"binary\n" + rows.tobytes())
    assert np.array_equal(back(p).view(np.uint32), pts.view(np.uint32))
    assert subprocess.call([exe, str(tmp_path / "missing.pcd"), str(tmp_path / "o.pcd")]) == 2


def test_pcd_corrupt_files_are_rejected(tmp_path):
    """Truncated / lying headers must fail like an unreadable file (exit 2), not allocate gigabytes or read out of bounds."""
    exe = str(tmp_path / "pcd_rt")
    (tmp_path / "rt.cpp").write_text(SRC)
    subprocess.check_call(["g++", "-O1", "-std=c++17", f"-I{ROOT}/include", f"-I{PKG}/host", "-o", exe, str(tmp_path / "rt.cpp")])
    n = 64
    pts = np.arange(n * 4, dtype=np.float32).reshape(n, 4)
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n"
           "WIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA ")
    soa = b"".join(pts[:, k].tobytes() for k in range(4))
    comp = lzf_literal(soa)
    cases = {
        "huge_points.pcd": hdr.format(n=10 ** 12).encode() + b"binary\n" + pts.tobytes(),
        "truncated_bin.pcd": hdr.format(n=n).encode() + b"binary\n" + pts.tobytes()[:100],
        "comp_no_sizes.pcd": hdr.format(n=n).encode() + b"binary_compressed\n" + b"\x01\x02",
        "comp_huge.pcd": hdr.format(n=n).encode() + b"binary_compressed\n" + struct.pack("<II", 0xFFFFFFF0, len(soa)) + comp,
        "comp_truncated.pcd": hdr.format(n=n).encode() + b"binary_compressed\n" + struct.pack("<II", len(comp), len(soa)) + comp[:50],
        # a back reference that points before the start of the output buffer
        "comp_bad_ref.pcd": hdr.format(n=n).encode() + b"binary_compressed\n" + struct.pack("<II", 2, len(soa)) + bytes([0x3F, 0xFF]),
        "ascii_short.pcd": hdr.format(n=n).encode() + b"ascii\n" + b"1 2 3 4\n",
    }
    for name, data in cases.items():
        p = str(tmp_path / name)
        open(p, "wb").write(data)
        assert subprocess.call([exe, p, str(tmp_path / "o.pcd")]) == 2, name
