"""Host-side sharding of the path across ranks (one process per GPU).

Stage A (per-map feature pipeline, map_merge_3d/src/map_merging.cpp:212-242) shards by contiguous blocks of maps;
stage B (the pair loop, :256-269) deals the row-major pair list out by LPT on an estimated pair cost.  Results are
written into their row-major slot so the pose graph sees the reference's pair order (Kruskal's tie-breaking depends on
it, src/graph.cpp:124).  Pure host logic: it is exercised on CPU with the gloo backend (tests/test_multi_rank.py) and
on GPUs with NCCL (bench.py).
"""
from __future__ import annotations

import numpy as np


def map_block(rank: int, world: int, n_maps: int):
    """(first, count, per) — the contiguous block of maps rank owns; `per` = block stride."""
    per = -(-n_maps // world)
    first = min(rank * per, n_maps)
    return first, max(0, min(per, n_maps - first)), per


def owner_of_map(m: int, world: int, n_maps: int) -> int:
    per = -(-n_maps // world)
    return min(m // per, world - 1)


def pair_list(n_keypoints):
    """Row-major i<j pairs of maps that both have keypoints (map_merging.cpp:246-254)."""
    m = len(n_keypoints)
    return [(i, j) for i in range(m - 1) for j in range(i + 1, m) if n_keypoints[i] > 0 and n_keypoints[j] > 0]


def pair_costs(ij, n_points, n_keypoints, dim):
    return [float(n_keypoints[i]) * n_keypoints[j] * dim * 2e-3 + 4.0 * n_points[i] + n_points[j] for i, j in ij]


def lpt_assign(costs, n_bins: int, ij=None):
    """Who registers which pair (Python restatement of csrc/dist.cu::lpt_assign; deterministic, identical on every rank).
    Pairs are dealt out in chunks that share a TARGET map (a rank builds the neighbour index and reach grid of every target
    its pairs name): the pairs of one target in ascending source order, cut into chunks of at most a quarter of a rank's
    fair share of the total cost, chunks assigned longest-processing-time-first.  Without `ij` every pair is its own chunk."""
    costs = np.asarray(costs, np.float64)
    owner = np.zeros(len(costs), np.int64)
    if len(costs) == 0 or n_bins <= 1:
        return owner
    if ij is None:
        chunks = [([k], float(costs[k])) for k in range(len(costs))]
    else:
        cap = float(costs.sum()) / n_bins / 4.0
        by_target = {}
        for k, (_a, b) in enumerate(ij):
            by_target.setdefault(int(b), []).append(k)
        chunks = []
        for b in sorted(by_target):
            v = by_target[b]
            i = 0
            while i < len(v):
                ids, c = [], 0.0
                while True:
                    c += float(costs[v[i]]); ids.append(v[i]); i += 1
                    if not (i < len(v) and c + float(costs[v[i]]) <= cap):
                        break
                chunks.append((ids, c))
    order = np.argsort(-np.array([c for _, c in chunks]), kind="stable")
    load = np.zeros(n_bins)
    for ci in order:
        ids, c = chunks[ci]
        b = int(np.argmin(load))
        load[b] += c
        owner[ids] = b
    return owner


def gather_pair_results(dist, torch, device, n_pairs: int, mine, T, conf):
    """All ranks contribute their pairs' (4x4 transform, confidence); returns float64 [n_pairs, 18] on every rank
    (column 17 counts contributions: exactly 1 everywhere when the shards are disjoint and complete)."""
    res = torch.zeros((n_pairs, 18), dtype=torch.float64, device=device)
    if len(mine):
        block = np.concatenate([np.asarray(T, np.float64).reshape(len(mine), 16), np.asarray(conf, np.float64).reshape(-1, 1),
                                np.ones((len(mine), 1))], axis=1)
        res[torch.as_tensor(list(mine), device=device)] = torch.from_numpy(block).to(device)
    if n_pairs and dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(res)  # disjoint slots: the sum is a gather
    return res


# ---- composeMaps sharded over ranks (map_merging.cpp:277-305) ------------------------------------------

def choose_splitters(hist, world: int):
    """Balanced key-range splitters from the all-reduced bucket histogram: int32[world + 1], splitters[r] <= bucket <
    splitters[r + 1] goes to rank r.  Deterministic, identical on every rank."""
    hist = np.asarray(hist, np.int64)
    nb = len(hist)
    cum = np.cumsum(hist)
    total = int(cum[-1]) if nb else 0
    sp = np.zeros(world + 1, np.int32)
    sp[world] = nb
    for r in range(1, world):
        target = total * r // world
        sp[r] = max(int(np.searchsorted(cum, target, side="left")) + (1 if total else 0), int(sp[r - 1]))
        sp[r] = min(int(sp[r]), nb)
    return sp


class CtxShardOps:
    """The per-rank steps of the sharded composeMaps on the CUDA library (include/mm3d.h, mm3d_compose_shard_*)."""

    def __init__(self, ctx, torch, device):
        self.ctx, self.torch, self.device = ctx, torch, device

    def begin(self, clouds, transforms):
        bbox, self.shard, n = self.ctx.compose_shard_begin(clouds, transforms)
        return bbox, n

    def histogram(self, gbbox, resolution, n_buckets):
        return self.ctx.compose_shard_histogram(self.shard, gbbox, resolution, n_buckets)

    def partition(self, gbbox, resolution, n_buckets, splitters, n):
        send = self.torch.empty((max(n, 1), 4), dtype=self.torch.float32, device=self.device)
        counts = self.ctx.compose_shard_partition(self.shard, gbbox, resolution, splitters, send.data_ptr(), n_buckets)
        return send[:n], counts

    def passthrough(self, n):
        """all of this rank's transformed points, unfiltered (pcl::VoxelGrid's overflow guard fired)"""
        return self.ctx.compose_shard_points(self.shard)

    def downsample(self, recv, resolution):
        self.torch.cuda.synchronize(self.device)
        return self.ctx.downsample_dev(recv.data_ptr(), recv.shape[0], resolution)

    def end(self):
        self.ctx.shard_free(self.shard)


def compose_sharded(ops, dist, torch, device, clouds_local, transforms_local, resolution, n_buckets: int = 4096):
    """composeMaps with the maps spread over ranks.  Returns this rank's slice of the composed map (float32 [n, 4]); the
    slices concatenated in rank order are bit-identical to the unsharded result.  The data path has exactly one
    exchange step — an all-to-all of raw points by owning key range — preceded by two small all-reduces (bounding box,
    key histogram)."""
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    bbox, n = ops.begin(clouds_local, transforms_local)
    lo = torch.tensor(bbox[:3], dtype=torch.float32, device=device)
    hi = torch.tensor(bbox[3:], dtype=torch.float32, device=device)
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    gbbox = np.concatenate([lo.cpu().numpy(), hi.cpu().numpy()]).astype(np.float32)
    total = torch.tensor([n], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(total)
    try:
        if int(total.item()) == 0:
            return np.zeros((0, 4), np.float32)
        hist = ops.histogram(gbbox, resolution, n_buckets)
        if hist is None:
            return ops.passthrough(n)  # rank order = map order: the concatenation is the reference's output
        h = torch.from_numpy(hist.astype(np.int64)).to(device)
        if world > 1:
            dist.all_reduce(h)
        splitters = choose_splitters(h.cpu().numpy(), world)
        send, counts = ops.partition(gbbox, resolution, n_buckets, splitters, n)
        if world == 1:
            return ops.downsample(send, resolution)
        sc = torch.from_numpy(np.asarray(counts, np.int64)).to(device)
        rc = torch.empty_like(sc)
        dist.all_to_all_single(rc, sc)
        recv_counts = [int(x) for x in rc.cpu().numpy()]
        recv = torch.empty((sum(recv_counts), 4), dtype=torch.float32, device=device)
        dist.all_to_all_single(recv, send.contiguous(), output_split_sizes=recv_counts, input_split_sizes=[int(x) for x in counts])
        return ops.downsample(recv, resolution)
    finally:
        ops.end()
