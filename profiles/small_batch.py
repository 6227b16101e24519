"""Fixed overhead of the two halves of the path at small batch sizes (what a rank sees at N = 4 / 8):
wall time vs. summed kernel time of features_compute on 1 / 2 / 4 / 8 maps and register_pairs on 3 / 7 / 14 / 28 pairs."""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import mm3d_pkg

mm = mm3d_pkg.load()
synth = importlib.import_module("map_merge_b200.synth")
sh = importlib.import_module("map_merge_b200.sharding")
maps, _ = synth.make_maps(**synth.CONFIGS["c2"])
ctx = mm.Context(0)
p = mm.default_params(descriptor_type="FPFH")
up = ctx.maps_upload(maps)


def timed(fn, reps=5):
    fn(); fn()
    torch.cuda.synchronize()
    ctx.profile_begin()
    l0 = ctx.launches
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / reps
    prof = ctx.profile_end()
    return wall, sum(k["ms"] for k in prof) / reps, (ctx.launches - l0) // reps, r


for count in (1, 2, 4, 8):
    wall, kern, nl, f = timed(lambda: ctx.features_compute(up, 0, count, p))
    print(f"features_compute on {count} map(s): wall {wall:.2f} ms, kernels {kern:.2f} ms, {nl} launches, gap {wall - kern:.2f} ms")
f = ctx.features_compute(up, 0, 8, p)
npt, nk, dim = f.sizes()
ij = sh.pair_list([int(x) for x in nk])
for n in (3, 7, 14, 28):
    wall, kern, nl, _ = timed(lambda: ctx.register_pairs(f, ij[:n], p))
    print(f"register_pairs on {n} pair(s): wall {wall:.2f} ms, kernels {kern:.2f} ms, {nl} launches, gap {wall - kern:.2f} ms")
