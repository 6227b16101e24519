"""How many columns of the descriptor distance matrix fall inside the tensor-core filter's error margin?  Runs the feature
pipeline on two maps of a config on the GPU, then analyses the FPFH descriptors on the host (numpy, float64)."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mm3d_pkg

mm = mm3d_pkg.load()
synth = importlib.import_module("map_merge_b200.synth")
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
maps, _ = synth.make_maps(**synth.CONFIGS[name])
maps = maps[:2]
ctx = mm.Context(0)
p = mm.default_params(descriptor_type=sys.argv[2] if len(sys.argv) > 2 else "FPFH")
up = ctx.maps_upload(maps)
f = ctx.features_compute(up, 0, 2, p)
descs = [f.export_host(i)[2] for i in range(2)]
a, b = [np.asarray(d, np.float64) for d in descs]
print("descriptors:", a.shape, b.shape)
ua = np.unique(a.astype(np.float32), axis=0); ub = np.unique(b.astype(np.float32), axis=0)
print("unique rows:", len(ua), len(ub))
mu = np.concatenate([a, b]).mean(0)
for centred in (False, True):
    A = a - mu if centred else a
    B = b - mu if centred else b
    na = (A * A).sum(1); nb = (B * B).sum(1)
    rng = np.random.default_rng(0)
    rows = rng.choice(len(A), 400, replace=False)
    Dm = na[rows, None] + nb[None, :] - 2 * A[rows] @ B.T
    kth = np.sort(Dm, axis=1)[:, 4]
    print(f"centred={centred}: median ||a||^2 {np.median(na):.0f}, max ||b||^2 {nb.max():.0f}, median 5th-NN distance {np.median(kth):.2f}, "
          f"5th-NN == 0 for {np.mean(kth < 1e-9) * 100:.1f}% of rows")
    for E in (3e-5, 1e-5, 3e-6):
        glob = 2 * E * (na[rows, None] + nb.max())
        loc = 2 * E * (na[rows, None] + nb[None, :])
        cg = (Dm <= kth[:, None] + glob).sum(1); cl = (Dm <= kth[:, None] + loc).sum(1)
        print(f"   E={E:g}: columns within margin (global bmax) mean {cg.mean():.1f} median {np.median(cg):.0f} p90 {np.percentile(cg, 90):.0f} max {cg.max()}"
              f" rows>32: {np.mean(cg > 32) * 100:.0f}%; (per-column norm) mean {cl.mean():.1f} median {np.median(cl):.0f} p90 {np.percentile(cl, 90):.0f}"
              f" max {cl.max()} rows>32: {np.mean(cl > 32) * 100:.0f}%")
m0 = ctx.knn_stats()
ctx.match(descs[0], descs[1], 5)
m1 = ctx.knn_stats()
print("kernel: rows", m1["rows"] - m0["rows"], "early flushes", m1["overflow_rows"] - m0["overflow_rows"], "exact evaluations/row",
      (m1["candidates"] - m0["candidates"]) / max(1, m1["rows"] - m0["rows"]))
