"""BASELINE-sized inputs on the GPU.

Against the CPU oracle (all host threads) wherever it finishes in seconds: maps 0-1 of configs 2 and 3, config 1 in full
through map_merge_tool, config 4 in miniature (2 x 500k, Harris + SHOT, inlier_threshold 0.2), composeMaps of 4 x 2M
points; the tensor-core k-NN against the exact scan on every pair of config 2 and 66 pairs of config 3.
Through size-independent properties where the oracle would take minutes (all 28 pairs of config 2): ground-truth
recovery, cycle consistency, idempotence and point conservation of the voxel grid, agreement between the resident path,
the host-buffer path and a sharded run."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import rot_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "map-merge_b200")


def _bits_equal(a, b):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32 if a.dtype.itemsize == 4 else np.uint64),
                                                 b.view(np.uint32 if b.dtype.itemsize == 4 else np.uint64))


def _pairs_against_oracle(ctx, oracle, maps, p_gpu, p_cpu, what):
    """Whole path on `maps`: per-pair transform and confidence bit for bit, correspondence / inlier counts equal, the stage
    outputs the oracle exposes (filtered cloud, keypoints, descriptors) bit for bit, global transforms within 1e-5."""
    oracle.set_threads(0)
    want = oracle.estimate_maps_transforms(maps, p_cpu)
    dm = ctx.maps_upload(maps)
    f = ctx.features_compute(dm, 0, len(maps), p_gpu)
    npt, nk, dim = f.sizes()
    ij = np.array([(i, j) for i in range(len(maps) - 1) for j in range(i + 1, len(maps)) if nk[i] > 0 and nk[j] > 0], np.int32).reshape(-1, 2)
    assert ij.tolist() == want["pairs"][:, :2].tolist(), what
    T, conf, stats = ctx.register_pairs(f, ij, p_gpu)
    assert stats[:, 0].tolist() == want["pairs"][:, 2].tolist(), f"{what}: correspondence counts {stats[:, 0]} vs {want['pairs'][:, 2]}"
    assert stats[:, 1].tolist() == want["pairs"][:, 3].tolist(), f"{what}: RANSAC inlier counts {stats[:, 1]} vs {want['pairs'][:, 3]}"
    assert _bits_equal(T, want["pair_T"]), f"{what}: pairwise transforms differ\n{T}\n{want['pair_T']}"
    assert _bits_equal(conf, want["pair_conf"]), f"{what}: confidences {conf} vs {want['pair_conf']}"
    G = ctx.estimate_resident(dm, p_gpu)
    np.testing.assert_allclose(G, want["transforms"], rtol=0, atol=1e-5)
    out = dict(features=f, n_points=npt, n_keypoints=nk, stats=stats, T=T)
    dm.free()
    return out


def test_config2_maps01_against_oracle(ctx, mm, oracle, synth):
    """configs[1] (8 x 500k, FPFH): maps 0-1 through the whole path against the oracle, every stage output bit for bit."""
    import oracle_py
    cfg = dict(synth.CONFIGS["c2"])
    maps, _ = synth.make_maps(**cfg, only=[0, 1])
    r = _pairs_against_oracle(ctx, oracle, maps[:2], mm.default_params(descriptor_type="FPFH"), oracle_py.default_params(descriptor_type=2), "c2")
    assert (r["n_points"] > 50_000).all() and (r["n_keypoints"] > 500).all() and r["stats"][0, 0] > 100
    # stage outputs of map 0
    p = oracle_py.default_params(descriptor_type=2)
    ds, _ = oracle.downsample(maps[0], p.resolution)
    fo, _, _ = oracle.remove_outliers(ds, p.descriptor_radius, p.outliers_min_neighbours)
    nm = oracle.normals(fo, p.normal_radius)
    kp = oracle.sift(fo, p.resolution, p.keypoint_threshold)
    kp, desc = oracle.fpfh(fo, nm, kp, p.descriptor_radius)
    pts, gkp, gdesc = r["features"].export_host(0)
    assert _bits_equal(pts, fo) and _bits_equal(gkp, kp) and _bits_equal(gdesc, desc)
    r["features"].free()


def test_config3_maps01_against_oracle(ctx, mm, oracle, synth):
    """configs[2] (32 x 1M, FPFH, the headline workload): maps 0-1 against the oracle."""
    import oracle_py
    cfg = dict(synth.CONFIGS["c3"])
    maps, _ = synth.make_maps(**cfg, only=[0, 1])
    r = _pairs_against_oracle(ctx, oracle, maps[:2], mm.default_params(descriptor_type="FPFH"), oracle_py.default_params(descriptor_type=2), "c3")
    assert (r["n_points"] > 150_000).all() and (r["n_keypoints"] > 2000).all() and r["stats"][0, 0] > 500
    r["features"].free()


def test_config4_mini_against_oracle(ctx, mm, oracle, synth):
    """configs[3] in miniature: 2 x 500k points, Harris3D + SHOT-1344, inlier_threshold 0.2."""
    import oracle_py
    cfg = dict(synth.CONFIGS["c4"])
    maps, _ = synth.make_maps(**cfg, only=[0, 1])
    pg = mm.default_params(keypoint_type="HARRIS", keypoint_threshold=0.0, descriptor_type="SHOT", inlier_threshold=0.2)
    pc = oracle_py.default_params(keypoint_type=1, keypoint_threshold=0.0, descriptor_type=4, inlier_threshold=0.2)
    r = _pairs_against_oracle(ctx, oracle, maps[:2], pg, pc, "c4")
    assert (r["n_keypoints"] > 20).all()
    r["features"].free()


def write_pcd(path, pts):
    n = len(pts)
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z rgb\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n"
           f"WIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA binary\n")
    with open(path, "wb") as f:
        f.write(hdr.encode())
        f.write(np.ascontiguousarray(pts, np.float32).tobytes())


def test_config1_cli_against_oracle(tmp_path, ctx, mm, oracle, synth):
    """configs[0] in full: map_merge_tool on two 200k-point room scans (SIFT + FPFH, RANSAC + ICP) against the oracle: the
    printed transforms to print precision, output.pcd bit for bit."""
    import oracle_py
    subprocess.check_call(["make", "-C", PKG, "tools"], stdout=subprocess.DEVNULL)
    maps, _ = synth.make_maps(**synth.CONFIGS["c1"])
    a, b = str(tmp_path / "a.pcd"), str(tmp_path / "b.pcd")
    write_pcd(a, maps[0]); write_pcd(b, maps[1])
    r = subprocess.run([os.path.join(PKG, "map_merge_tool"), a, b, "--descriptor_type", "FPFH"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    nums = re.findall(r"[-+]?\d*\.?\d+(?:[eE][-+]?\d+)?", r.stdout.split("Estimated transforms:")[1].split("> Compositing")[0])
    T = np.array(nums, np.float64).reshape(-1, 4, 4)
    oracle.set_threads(0)
    want = oracle.estimate_maps_transforms(maps, oracle_py.default_params(descriptor_type=2))["transforms"]
    np.testing.assert_allclose(T, want, atol=2e-5, rtol=1e-5)  # printed with 6 significant digits
    raw = open(tmp_path / "output.pcd", "rb").read()
    k = raw.index(b"DATA binary\n") + len(b"DATA binary\n")
    out = np.frombuffer(raw[k:], np.float32).reshape(-1, 4)
    # the tool composes with ITS transforms (within 1e-5 of the oracle's: general 4x4 inverses in the graph step); the same
    # call through the C ABI returns them with all their bits, and the oracle composes with exactly those
    T_gpu = ctx.estimate_maps_transforms(maps, mm.default_params(descriptor_type="FPFH"))
    np.testing.assert_allclose(T, T_gpu, atol=2e-5, rtol=1e-5)
    comp = oracle.compose_maps(maps, T_gpu, 0.05)
    assert _bits_equal(out, comp)


def test_compose_large_against_oracle(ctx, oracle, synth):
    """composeMaps of 4 x 2M points at output_resolution 0.05 (config 5 in miniature) against the oracle, bit for bit; one map
    carries a zero transform and is skipped (map_merging.cpp:293-295)."""
    maps, truth = synth.make_maps(seed=5, n_maps=4, n_points=2_000_000, size_x=28.0, size_y=20.0, rooms_x=2, rooms_y=2)
    T = np.stack([np.linalg.inv(truth[0]) @ t for t in truth]).astype(np.float32)
    got = ctx.compose_maps(maps, T, 0.05)
    want = oracle.compose_maps(maps, T, 0.05)
    assert _bits_equal(got, want)
    T[2] = 0
    assert _bits_equal(ctx.compose_maps(maps, T, 0.05), oracle.compose_maps(maps, T, 0.05))


def _knn_modes(ctx, monkeypatch, a, b, k=5):
    monkeypatch.setenv("MM3D_KNN", "exact")
    ie, de = ctx.knn(a, b, k)
    monkeypatch.setenv("MM3D_KNN", "tc")
    it, dt = ctx.knn(a, b, k)
    monkeypatch.delenv("MM3D_KNN")
    return ie, de, it, dt


def _seq_fp32_dist(a, b):
    """flann::L2_Simple as the exact scan evaluates it: float32, dimension by dimension, no contraction."""
    acc = np.zeros((len(a), len(b)), np.float32)
    for t in range(a.shape[1]):
        diff = a[:, None, t] - b[None, :, t]
        acc += diff * diff
    return acc


def test_tensor_core_knn_error_bound(ctx, mm, synth):
    """The filter's error model on the headline workload's descriptors (two c3 maps, both directions):
    v = acc + (1 - ES) ||a'||^2 must satisfy v <= d <= v + 2 ES (||a'||^2 + ||b'||^2) for EVERY pair of rows, d = the sequential
    FP32 distance of the exact scan — with room to spare: the measured deviation of the dot product must stay below ES / 2."""
    cfg = dict(synth.CONFIGS["c3"])
    maps, _ = synth.make_maps(**cfg, only=[0, 1])
    p = mm.default_params(descriptor_type="FPFH")
    dm = ctx.maps_upload(maps[:2])
    f = ctx.features_compute(dm, 0, 2, p)
    desc = [f.export_host(m)[2] for m in range(2)]
    f.free(); dm.free()
    rng = np.random.default_rng(0)
    worst = 0.0
    for a, b in ((desc[0], desc[1]), (desc[1], desc[0])):
        a = a[np.sort(rng.choice(len(a), 2048, replace=False))]
        au = ctx.knn_tc_audit(a, b, 5)
        es = au["err_store"]
        d = _seq_fp32_dist(a, b).astype(np.float64)
        na = au["norm_a"].astype(np.float64)[:, None]; nb = au["norm_b"].astype(np.float64)[None, :]
        v = au["acc"].astype(np.float64) + (1.0 - es) * na
        ratio = (d - v) / (na + nb + 1e-30)      # ES +- (deviation of the dot product) / (||a'||^2 + ||b'||^2)
        assert ratio.min() >= 0.0, f"lower bound violated: min ratio {ratio.min():.3e}"
        assert ratio.max() <= 2.0 * es, f"upper bound violated: max ratio {ratio.max():.3e} vs {2 * es:.3e}"
        dev = np.abs(ratio - es).max()
        worst = max(worst, dev)
        assert dev < 0.5 * es, f"deviation {dev:.3e} leaves less than half of ES = {es:.3e} as margin"
        # and the search itself equals numpy's (distance, index) ranking
        order = np.lexsort((np.broadcast_to(np.arange(d.shape[1]), d.shape), d), axis=1)[:, :5]
        assert np.array_equal(au["idx"], order)
    print(f"tensor-core filter: worst deviation / (||a'||^2 + ||b'||^2) = {worst:.3e}, ES = {es:.3e}")


def test_tensor_core_knn_against_exact_scan_config2(ctx, mm, synth, monkeypatch):
    """The tcgen05 filter + exact re-rank against the FP32 scan where its error bound is under stress: the clustered FPFH
    descriptors of config 2 — indices and distance bits of both directions of all 28 pairs."""
    maps, _ = synth.make_maps(**synth.CONFIGS["c2"])
    p = mm.default_params(descriptor_type="FPFH")
    dm = ctx.maps_upload(maps)
    f = ctx.features_compute(dm, 0, len(maps), p)
    desc = [f.export_host(m)[2] for m in range(len(maps))]
    f.free(); dm.free()
    rows = 0
    for i in range(len(maps) - 1):
        for j in range(i + 1, len(maps)):
            for a, b in ((desc[i], desc[j]), (desc[j], desc[i])):
                ie, de, it, dt = _knn_modes(ctx, monkeypatch, a, b)
                assert np.array_equal(ie, it), f"pair ({i}, {j}): {int((ie != it).any(axis=1).sum())} rows with different neighbours"
                assert _bits_equal(de, dt), f"pair ({i}, {j}): distances differ"
                rows += len(a)
    assert rows > 28 * 2 * 500


def test_tensor_core_knn_against_exact_scan_config3(ctx, mm, synth, monkeypatch):
    """The same on the headline workload: the first 12 maps of config 3, 66 pairs, both directions."""
    cfg = dict(synth.CONFIGS["c3"])
    maps, _ = synth.make_maps(**cfg, only=range(12))
    p = mm.default_params(descriptor_type="FPFH")
    dm = ctx.maps_upload(maps[:12])
    f = ctx.features_compute(dm, 0, 12, p)
    desc = [f.export_host(m)[2] for m in range(12)]
    f.free(); dm.free()
    n_pairs = 0
    for i in range(11):
        for j in range(i + 1, 12):
            for a, b in ((desc[i], desc[j]), (desc[j], desc[i])):
                ie, de, it, dt = _knn_modes(ctx, monkeypatch, a, b)
                assert np.array_equal(ie, it), f"pair ({i}, {j}): {int((ie != it).any(axis=1).sum())} rows with different neighbours"
                assert _bits_equal(de, dt), f"pair ({i}, {j}): distances differ"
            n_pairs += 1
    assert n_pairs == 66


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
