import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def mm():
    # the built libraries are not in the history: build them when a fresh checkout runs the tests before build()
    if not os.path.exists(os.path.join(ROOT, "map-merge_b200", "libmm3d.so")) or \
            not os.path.exists(os.path.join(ROOT, "map-merge_b200", "libmm3d_shim.so")):
        import __graft_entry__
        __graft_entry__.build()
    import mm3d_pkg
    return mm3d_pkg.load()


@pytest.fixture(scope="session")
def synth():
    import mm3d_pkg
    return mm3d_pkg.load_synth()


@pytest.fixture(scope="session")
def oracle():
    import oracle_py
    return oracle_py.Oracle()


@pytest.fixture(scope="session")
def oracle_libm():
    import oracle_py
    return oracle_py.Oracle(libm=True)


@pytest.fixture(scope="session")
def ctx(mm):
    c = mm.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def tiny_maps(synth):
    return synth.make_maps(**synth.CONFIGS["tiny"])


@pytest.fixture(scope="session")
def small_maps(synth):
    return synth.make_maps(**synth.CONFIGS["small"])


@pytest.fixture(scope="session")
def tiny_stages(oracle, tiny_maps):
    """Oracle outputs of every per-map stage for the two 'tiny' maps (inputs of the stage-level parity tests)."""
    maps, _ = tiny_maps
    out = []
    for m in maps:
        ds, meta = oracle.downsample(m, 0.1)
        fo, kept, cnt = oracle.remove_outliers(ds, 0.8, 50)
        nm = oracle.normals(fo, 0.6)
        kp, dog, sc = oracle.sift(fo, 0.1, 5.0, debug=True)
        kp2, desc, spfh = oracle.fpfh(fo, nm, kp, 0.8, debug=True)
        out.append(dict(raw=m, ds=ds, meta=meta, filtered=fo, kept=kept, counts=cnt, normals=nm, kp_sift=kp, dog0=dog, kp=kp2, desc=desc,
                        spfh=spfh))
    return out


def rot_err(A, B):
    R = A[:3, :3].astype(np.float64) @ B[:3, :3].astype(np.float64).T
    return float(np.arccos(np.clip((np.trace(R) - 1) / 2, -1, 1)))
