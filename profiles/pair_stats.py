"""Per-pair statistics of the c3 workload (first 12 maps, 66 pairs): correspondences, RANSAC inliers, ICP iterations."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mm3d_pkg
mm = mm3d_pkg.load(); synth = mm3d_pkg.load_synth()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
maps, _ = synth.make_maps(**synth.CONFIGS["c3"], only=range(n))
ctx = mm.Context(0)
p = mm.default_params(descriptor_type="FPFH")
dm = ctx.maps_upload(maps[:n])
f = ctx.features_compute(dm, 0, n, p)
ij = np.array([(i, j) for i in range(n - 1) for j in range(i + 1, n)], np.int32)
for _ in range(2):
    t0 = time.time(); T, conf, stats = ctx.register_pairs(f, ij, p); dt = time.time() - t0
print(f"{len(ij)} pairs in {dt * 1e3:.1f} ms")
it = stats[:, 2]
print("ICP iterations: histogram", np.bincount(it, minlength=20).tolist())
print("correspondences: min/median/max", stats[:, 0].min(), int(np.median(stats[:, 0])), stats[:, 0].max())
print("inliers: min/median/max", stats[:, 1].min(), int(np.median(stats[:, 1])), stats[:, 1].max(), " failed RANSAC (0 inliers):", int((stats[:, 1] == 0).sum()))
slow = np.argsort(-it)[:8]
for k in slow:
    print(f"  pair {ij[k].tolist()}: iterations {it[k]}, converged {stats[k, 3]}, corr {stats[k, 0]}, inliers {stats[k, 1]}, conf {conf[k]:.2f}")
