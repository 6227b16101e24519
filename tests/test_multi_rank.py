"""The N > 1 host logic on CPU: two gloo ranks shard maps and pairs, exchange results, rank 0 builds the pose graph.
The per-pair registration itself is GPU work (covered by the -m gpu tests); here each rank looks its pairs up in a
precomputed table so that sharding, ordering and the result exchange are what is under test."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _poses(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        q, _r = np.linalg.qr(rng.normal(size=(3, 3)))
        if np.linalg.det(q) < 0:
            q[:, 0] *= -1
        P = np.eye(4)
        P[:3, :3] = q
        P[:3, 3] = rng.normal(size=3) * 4
        out.append(P)
    return out


def _table(n_maps, seed):
    poses = _poses(n_maps, seed)
    rng = np.random.default_rng(seed + 1)
    n_kp = rng.integers(0, 5000, n_maps)
    n_kp[rng.integers(0, n_maps)] = 0  # one map without keypoints drops out of the pair list
    n_pts = rng.integers(50_000, 150_000, n_maps)
    tab = {}
    for i in range(n_maps - 1):
        for j in range(i + 1, n_maps):
            tab[(i, j)] = ((np.linalg.inv(poses[j]) @ poses[i]).astype(np.float32), float(np.round(rng.uniform(0.5, 20.0), 1)))
    return n_kp, n_pts, tab


def _worker(rank, world, port, n_maps, seed, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import mm3d_pkg
    import importlib
    mm = mm3d_pkg.load()
    sh = importlib.import_module("map_merge_b200.sharding")
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        n_kp, n_pts, tab = _table(n_maps, seed)
        first, count, per = sh.map_block(rank, world, n_maps)
        # stage A exchange: every rank contributes the sizes of its block of maps
        sizes = torch.zeros((per, 2), dtype=torch.int32)
        for m in range(count):
            sizes[m, 0] = int(n_pts[first + m]); sizes[m, 1] = int(n_kp[first + m])
        allsz = [torch.zeros_like(sizes) for _ in range(world)]
        dist.all_gather(allsz, sizes)
        allsz = torch.stack(allsz).numpy()
        g_pts = [int(allsz[sh.owner_of_map(m, world, n_maps), m - sh.owner_of_map(m, world, n_maps) * per, 0]) for m in range(n_maps)]
        g_kp = [int(allsz[sh.owner_of_map(m, world, n_maps), m - sh.owner_of_map(m, world, n_maps) * per, 1]) for m in range(n_maps)]
        assert g_pts == [int(x) for x in n_pts] and g_kp == [int(x) for x in n_kp]
        # stage B: shard the pair list, "register" the local share, exchange
        ij = sh.pair_list(g_kp)
        owner = sh.lpt_assign(sh.pair_costs(ij, g_pts, g_kp, 33), world, ij)
        mine = [k for k in range(len(ij)) if owner[k] == rank]
        T = np.stack([tab[ij[k]][0] for k in mine]) if mine else np.zeros((0, 4, 4), np.float32)
        conf = np.array([tab[ij[k]][1] for k in mine])
        res = sh.gather_pair_results(dist, torch, torch.device("cpu"), len(ij), mine, T.transpose(0, 2, 1), conf).numpy()
        if rank == 0:
            G, ref = mm.global_transforms(np.array(ij, np.int32), res[:, :16].reshape(-1, 4, 4).transpose(0, 2, 1).astype(np.float32), res[:, 16], 0.0)
            q.put(dict(ij=ij, res=res, G=G, ref=ref, mine=len(mine)))
        else:
            q.put(dict(mine=len(mine)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_maps,world", [(8, 2), (5, 2), (3, 2)])
def test_sharded_path_matches_single_process(mm, n_maps, world):
    import torch.multiprocessing as mp
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_maps, 7, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    r0 = [o for o in outs if "G" in o][0]
    import importlib
    sh = importlib.import_module("map_merge_b200.sharding")
    n_kp, n_pts, tab = _table(n_maps, 7)
    ij = sh.pair_list([int(x) for x in n_kp])
    assert r0["ij"] == ij
    assert sum(o["mine"] for o in outs) == len(ij)
    assert (r0["res"][:, 17] == 1).all()  # every pair registered by exactly one rank
    T = np.stack([tab[p][0] for p in ij]); conf = [tab[p][1] for p in ij]
    np.testing.assert_array_equal(r0["res"][:, :16].reshape(-1, 4, 4).transpose(0, 2, 1).astype(np.float32), T)
    G, ref = mm.global_transforms(np.array(ij, np.int32), T, conf, 0.0)
    assert ref == r0["ref"]
    np.testing.assert_array_equal(G, r0["G"])


def test_lpt_is_balanced_and_deterministic(mm):
    import importlib
    sh = importlib.import_module("map_merge_b200.sharding")
    rng = np.random.default_rng(0)
    costs = rng.uniform(1, 10, 496)
    a = sh.lpt_assign(costs, 8)
    assert np.array_equal(a, sh.lpt_assign(costs, 8))
    loads = np.array([costs[a == b].sum() for b in range(8)])
    assert loads.max() / loads.mean() < 1.02
    assert sh.map_block(0, 8, 32) == (0, 4, 4) and sh.map_block(7, 8, 32) == (28, 4, 4)
    assert sh.map_block(2, 4, 5) == (4, 1, 2) and sh.map_block(3, 4, 5)[1] == 0


# ---- composeMaps sharded over ranks -------------------------------------------------------------------------------------
class OracleShardOps:
    """The per-rank steps of sharding.compose_sharded restated on the CPU checker (numpy + oracle), so that the host logic
    — box / histogram all-reduces, splitter choice, the all-to-all — runs under gloo without a GPU."""

    def __init__(self, oracle, torch):
        self.o, self.torch = oracle, torch

    def begin(self, clouds, transforms):
        r = self.o.compose_maps(clouds, transforms, 0.0) if len(clouds) else None  # leaf 0 -> transformed concatenation
        self.pts = r if r is not None else np.zeros((0, 4), np.float32)
        big = np.float32(np.finfo(np.float32).max)
        if len(self.pts) == 0:
            return np.array([big, big, big, -big, -big, -big], np.float32), 0
        return np.concatenate([self.pts[:, :3].min(0), self.pts[:, :3].max(0)]).astype(np.float32), len(self.pts)

    def _buckets(self, gbbox, resolution, n_buckets):
        inv = np.float32(1.0) / np.float32(resolution)
        ext = [int((gbbox[3 + k] - gbbox[k]) * inv) + 1 for k in range(3)]
        if not resolution > 0 or ext[0] * ext[1] * ext[2] > 2**31 - 1:
            return None
        mn = [int(np.floor(gbbox[k] * inv)) for k in range(3)]
        dv = [int(np.floor(gbbox[3 + k] * inv)) - mn[k] + 1 for k in range(3)]
        width = max(1, -(-(dv[0] * dv[1] * dv[2]) // n_buckets))
        ijk = [(np.floor(self.pts[:, k] * inv) - np.float32(mn[k])).astype(np.int64) for k in range(3)]
        key = ijk[0] + ijk[1] * dv[0] + ijk[2] * dv[0] * dv[1]
        return np.minimum(key // width, n_buckets - 1)

    def histogram(self, gbbox, resolution, n_buckets):
        b = self._buckets(gbbox, resolution, n_buckets)
        return None if b is None else np.bincount(b, minlength=n_buckets).astype(np.uint64)

    def partition(self, gbbox, resolution, n_buckets, splitters, n):
        b = self._buckets(gbbox, resolution, n_buckets)
        dest = np.searchsorted(np.asarray(splitters[1:-1]), b, side="right")
        order = np.argsort(dest, kind="stable")
        return self.torch.from_numpy(self.pts[order].copy()), np.bincount(dest, minlength=len(splitters) - 1)

    def passthrough(self, n):
        return self.pts

    def downsample(self, recv, resolution):
        return self.o.downsample(recv.numpy(), resolution)[0] if recv.shape[0] else np.zeros((0, 4), np.float32)

    def end(self):
        pass


def _compose_case(seed, n_maps):
    sys.path.insert(0, ROOT)
    import mm3d_pkg
    import importlib
    mm3d_pkg.load()
    synth = importlib.import_module("map_merge_b200.synth")
    maps, truth = synth.make_maps(seed, n_maps, 6000, 20.0, 10.0, 2, 1)
    T = np.stack([np.linalg.inv(truth[0]) @ t for t in truth]).astype(np.float32)
    if n_maps > 2:
        T[2] = 0  # a map that could not be placed is skipped (map_merging.cpp:293-295)
    return maps, T


def _compose_worker(rank, world, port, n_maps, resolution, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    import importlib
    import oracle_py
    maps, T = _compose_case(5, n_maps)
    sh = importlib.import_module("map_merge_b200.sharding")
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        first, count, _per = sh.map_block(rank, world, n_maps)
        ops = OracleShardOps(oracle_py.Oracle(), torch)
        out = sh.compose_sharded(ops, dist, torch, torch.device("cpu"), maps[first:first + count], T[first:first + count], resolution)
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_maps,world,resolution", [(4, 2, 0.05), (3, 2, 0.2), (1, 2, 0.05), (4, 2, 1e-3)])
def test_sharded_compose_matches_unsharded(oracle, n_maps, world, resolution):
    import torch.multiprocessing as mp
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_compose_worker, args=(r, world, port, n_maps, resolution, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    maps, T = _compose_case(5, n_maps)
    want = oracle.compose_maps(maps, T, resolution)
    got = np.concatenate([outs[r] for r in range(world)])
    assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    if n_maps >= 3 and resolution > 1e-2:
        sizes = [len(outs[r]) for r in range(world)]
        assert min(sizes) > 0.3 * sum(sizes) / world  # the key-range splitters balance the ranks


def test_choose_splitters(mm):
    import importlib
    sh = importlib.import_module("map_merge_b200.sharding")
    h = np.zeros(64, np.int64); h[10:20] = 100
    sp = sh.choose_splitters(h, 4)
    assert sp[0] == 0 and sp[-1] == 64 and (np.diff(sp) >= 0).all()
    loads = [h[sp[r]:sp[r + 1]].sum() for r in range(4)]
    assert sum(loads) == 1000 and max(loads) <= 300
    assert list(sh.choose_splitters(np.zeros(8, np.int64), 2)) == [0, 0, 8]
    assert list(sh.choose_splitters(np.array([5]), 3)) == [0, 1, 1, 1]


def test_library_pair_plan_matches_python_restatement(mm):
    """mm3d_dist_block / mm3d_dist_plan (csrc/dist.cu, what every rank of the multi-GPU path computes for itself) against
    the Python restatement in sharding.py: same blocks, same row-major pair list, same LPT owners."""
    import importlib
    sh = importlib.import_module("map_merge_b200.sharding")
    rng = np.random.default_rng(3)
    for n_maps, world in ((32, 8), (32, 4), (8, 2), (5, 3), (2, 2), (1, 4), (7, 16)):
        n_pts = rng.integers(100_000, 300_000, n_maps)
        n_kp = rng.integers(0, 9000, n_maps)
        n_kp[rng.integers(0, n_maps)] = 0
        pairs, owner = mm.dist_plan(n_pts, n_kp, 33, world)
        ij = sh.pair_list(n_kp.tolist())
        assert pairs.tolist() == [list(p) for p in ij]
        want = sh.lpt_assign(sh.pair_costs(ij, n_pts.tolist(), n_kp.tolist(), 33), world, ij) if ij else np.zeros(0, np.int64)
        assert owner.tolist() == [int(x) for x in want]
        if len(ij) >= 4 * world:
            costs = np.array(sh.pair_costs(ij, n_pts.tolist(), n_kp.tolist(), 33))
            loads = np.array([costs[owner == r].sum() for r in range(world)])
            assert loads.max() <= loads.mean() * 1.25 + costs.max()  # chunks are at most a quarter of a fair share
            if n_maps == 32 and world == 8:
                # the point of chunking by target: a rank names few target maps (it builds their index and reach grid)
                targets = [len({ij[k][1] for k in range(len(ij)) if owner[k] == r}) for r in range(world)]
                assert max(targets) <= 12, targets
        covered = []
        for r in range(world):
            first, count = mm.dist_block(r, world, n_maps)
            assert (first, count) == sh.map_block(r, world, n_maps)[:2]
            covered += list(range(first, first + count))
        assert covered == list(range(n_maps))


def _settle_worker(rank, world, port, q):
    """bench.settle() with a step that is collective, as every N > 1 step is (NCCL inside the library)."""
    sys.path.insert(0, ROOT)
    import types
    import torch
    import torch.distributed as dist
    import bench
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        calls = []

        def step():
            t = torch.ones(1)
            dist.all_reduce(t)  # blocks until every rank has entered the step
            calls.append(int(t.item()))

        job = types.SimpleNamespace(rank=rank, world=world, torch=torch, dist=dist, dev="cpu")
        bench.settle(job, step, seconds=0.05)
        q.put({"rank": rank, "calls": len(calls), "sum_ok": all(c == world for c in calls)})
    finally:
        dist.destroy_process_group()


def test_bench_settle_is_collective_safe():
    """Round 2 regression: rank 0 used to run the first untimed step alone while the others waited for its broadcast —
    a deadlock at every N > 1 (found by the 8-GPU run, which sat in the NCCL watchdog for ten minutes)."""
    import torch.multiprocessing as mp
    import socket
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_settle_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        outs = [q.get(timeout=90) for _ in range(world)]
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.terminate()
    assert all(o["sum_ok"] for o in outs)
    assert outs[0]["calls"] == outs[1]["calls"] >= 1  # the same number of steps on every rank
