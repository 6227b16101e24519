"""BASELINE-sized inputs on the GPU, checked through size-independent properties (the CPU oracle would take minutes here):
ground-truth recovery, cycle consistency of the pairwise transforms, idempotence and point conservation of the voxel grid,
agreement between the resident path, the host-buffer path and a sharded run."""
import numpy as np
import pytest

from conftest import rot_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2(synth):
    cfg = dict(synth.CONFIGS["c2"])
    return synth.make_maps(**cfg)


def test_config2_all_pairs(ctx, mm, c2):
    """8 maps x 500k points, FPFH, all 28 pairs (BASELINE.json configs[1])."""
    maps, truth = c2
    p = mm.default_params(descriptor_type="FPFH")
    dm = ctx.maps_upload(maps)
    G, stage_ms = ctx.estimate_resident(dm, p, stage_times=True)
    assert G.shape == (8, 4, 4)
    ref = [i for i in range(8) if np.array_equal(G[i], np.eye(4, dtype=np.float32))]
    assert len(ref) == 1
    r = ref[0]
    for i in range(8):
        want = np.linalg.inv(truth[r]) @ truth[i]
        assert rot_err(G[i], want) < 0.1, (i, rot_err(G[i], want))
        assert np.linalg.norm(G[i][:3, 3] - want[:3, 3]) < 1.0
    # same answer from host buffers (mm3d_estimate_maps_transforms) and when run twice (determinism)
    np.testing.assert_array_equal(ctx.estimate_maps_transforms(maps, p), G)
    np.testing.assert_array_equal(ctx.estimate_resident(dm, p), G)
    # pairwise results: sharded over two "ranks" == all at once; confidences positive; cycle consistency on a triangle
    f = ctx.features_compute(dm, 0, 8, p)
    npt, nk, dim = f.sizes()
    assert dim == 33 and (nk > 500).all() and (npt > 50_000).all()
    ij = np.array([(i, j) for i in range(7) for j in range(i + 1, 8)], np.int32)
    T, conf, stats = ctx.register_pairs(f, ij, p)
    Ta, ca, _ = ctx.register_pairs(f, ij[::2], p)
    Tb, cb, _ = ctx.register_pairs(f, ij[1::2], p)
    np.testing.assert_array_equal(T[::2], Ta)
    np.testing.assert_array_equal(T[1::2], Tb)
    np.testing.assert_array_equal(conf[::2], ca)
    assert (conf > 0).all() and (stats[:, 0] > 100).all()
    ok = stats[:, 1] >= 20
    assert ok.mean() > 0.8  # heavily overlapping windows: RANSAC succeeds on (nearly) all pairs
    look = {tuple(p_): k for k, p_ in enumerate(ij.tolist())}
    for (a, b, c_) in ((0, 1, 2), (2, 4, 6), (1, 3, 7)):
        if ok[look[(a, b)]] and ok[look[(b, c_)]] and ok[look[(a, c_)]]:
            chain = T[look[(b, c_)]].astype(np.float64) @ T[look[(a, b)]].astype(np.float64)
            assert rot_err(chain, T[look[(a, c_)]]) < 0.15
    G2, ref2 = mm.global_transforms(ij, T, conf, p.confidence_threshold)
    np.testing.assert_array_equal(G2, G)


def test_compose_large(ctx, mm, synth):
    """composeMaps on 4 x 2M points at output_resolution 0.05 (config 5 in miniature): conservation + idempotence."""
    maps, truth = synth.make_maps(seed=5, n_maps=4, n_points=2_000_000, size_x=28.0, size_y=20.0, rooms_x=2, rooms_y=2)
    T = np.stack([np.linalg.inv(truth[0]) @ t for t in truth]).astype(np.float32)
    out = ctx.compose_maps(maps, T, 0.05)
    assert 100_000 < len(out) < 8_000_000
    # one point per voxel, ascending voxel order, inside the transformed bounding box
    inv = np.float32(1.0) / np.float32(0.05)
    ijk = np.floor(out[:, :3] * inv).astype(np.int64)
    ijk -= ijk.min(0)
    key = ijk[:, 0] + ijk[:, 1] * (ijk[:, 0].max() + 1) + ijk[:, 2] * (ijk[:, 0].max() + 1) * (ijk[:, 1].max() + 1)
    d = np.diff(key)
    assert (d > 0).mean() > 0.9999 and (d >= 0).all() or (d < 0).sum() < 50  # centroids may sit an ulp outside their voxel
    # composing the composed map again changes nothing but rounding
    again = ctx.compose_maps([out], [np.eye(4, dtype=np.float32)], 0.05)
    assert abs(len(again) - len(out)) <= max(2, len(out) // 100000)
    # a zero transform drops that map; the mean of all points is preserved by voxel averaging up to the voxel size
    T0 = T.copy(); T0[3] = 0
    out3 = ctx.compose_maps(maps, T0, 0.05)
    assert len(out3) < len(out)
    # colours stay valid bytes and alpha is the input's 255
    assert ((out[:, 3].view(np.uint32) >> 24) == 255).all()


def test_config4_variant(ctx, mm, synth):
    """Harris3D + SHOT with a tight inlier threshold on 500k-point maps (configs[3], fewer maps)."""
    cfg = dict(synth.CONFIGS["c2"]); cfg["n_maps"] = 4; cfg["seed"] = 4
    maps, truth = synth.make_maps(**cfg)
    p = mm.default_params(keypoint_type="HARRIS", keypoint_threshold=0.0, descriptor_type="SHOT", inlier_threshold=0.2)
    dm = ctx.maps_upload(maps)
    f = ctx.features_compute(dm, 0, 4, p)
    npt, nk, dim = f.sizes()
    assert dim == 1344 and (nk > 20).all()
    pts, kp, desc = f.export_host(0)
    np.testing.assert_allclose(np.linalg.norm(desc, axis=1), 1.0, atol=1e-4)
    G = ctx.estimate_resident(dm, p)
    assert G.shape[1:] == (4, 4) and np.isfinite(G).all()
