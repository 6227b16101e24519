"""Tensor-core k-NN vs the exact scan on two c3 maps: differing rows in detail, filter statistics, timing."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mm3d_pkg
mm = mm3d_pkg.load(); synth = mm3d_pkg.load_synth()
ma, mb = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (0, 4)
maps, _ = synth.make_maps(**synth.CONFIGS["c3"], only=[ma, mb])
ctx = mm.Context(0)
p = mm.default_params(descriptor_type="FPFH")
dm = ctx.maps_upload([maps[ma], maps[mb]])
f = ctx.features_compute(dm, 0, 2, p)
desc = [f.export_host(m)[2] for m in range(2)]
for a, b, tag in ((desc[0], desc[1], "a->b"), (desc[1], desc[0], "b->a")):
    os.environ["MM3D_KNN"] = "exact"
    ie, de = ctx.knn(a, b, 5)
    os.environ["MM3D_KNN"] = "tc"
    s0 = ctx.knn_stats()
    t0 = time.time(); it, dt = ctx.knn(a, b, 5); t1 = time.time()
    s1 = ctx.knn_stats()
    rows = s1["rows"] - s0["rows"]
    print(f"{tag}: {len(a)} x {len(b)}; tc call {1e3 * (t1 - t0):.1f} ms; rows {rows}, flushes/row {(s1['overflow_rows'] - s0['overflow_rows']) / max(rows, 1):.2f}, "
          f"exact evaluations/row {(s1['candidates'] - s0['candidates']) / max(rows, 1):.1f}")
    bad = np.where((ie != it).any(axis=1) | (de.view(np.uint32) != dt.view(np.uint32)).any(axis=1))[0]
    print(f"  rows that differ: {len(bad)}")
    for r in bad[:6]:
        print(f"  row {r}: exact idx {ie[r].tolist()} dist {de[r].tolist()}")
        print(f"          tc    idx {it[r].tolist()} dist {dt[r].tolist()}")
        # exact distances of the tc picks, by numpy
        acc = np.zeros(len(b), np.float32)
        for t in range(33):
            diff = a[r, t] - b[:, t]
            acc += diff * diff
        order = np.lexsort((np.arange(len(b)), acc))[:8]
        print(f"          numpy top-8 idx {order.tolist()} dist {acc[order].tolist()}")
    if len(bad):
        sub = a[bad[:256]]
        au = ctx.knn_tc_audit(sub, b, 5)
        es = au["err_store"]
        d = np.zeros((len(sub), len(b)), np.float32)
        for t in range(33):
            diff = sub[:, None, t] - b[None, :, t]
            d += diff * diff
        na = au["norm_a"].astype(np.float64)[:, None]; nb = au["norm_b"].astype(np.float64)[None, :]
        v = au["acc"].astype(np.float64) + (1.0 - es) * na
        ratio = (d.astype(np.float64) - v) / (na + nb + 1e-30)
        print(f"  audit of the differing rows: ratio min {ratio.min():.3e} max {ratio.max():.3e} (bounds 0 .. {2 * es:.3e}); "
              f"rows violating: {int(((ratio < 0) | (ratio > 2 * es)).any(axis=1).sum())}; audit-kernel idx equal to exact: "
              f"{int((au['idx'] == ie[bad[:256]]).all(axis=1).sum())} of {len(sub)}")
        r0 = 0
        order = np.lexsort((np.arange(len(b)), d[r0]))[:6]
        print(f"  row {bad[0]}: norms a {na[r0, 0]:.1f}; top-6 cols {order.tolist()} d {d[r0][order].tolist()} v {v[r0][order].tolist()} nb {nb[0][order].tolist()}")
