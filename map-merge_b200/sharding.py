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


def lpt_assign(costs, n_bins: int):
    """Longest-processing-time-first assignment; deterministic, identical on every rank."""
    order = np.argsort(-np.asarray(costs, np.float64), kind="stable")
    load = np.zeros(n_bins)
    owner = np.zeros(len(costs), np.int64)
    for k in order:
        b = int(np.argmin(load))
        owner[k] = b
        load[b] += costs[k]
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
