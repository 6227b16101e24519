"""Tensor-core k-NN vs the exact scan on the small golden match cases (ties, tiny sets, 2-D descriptors)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np
import mm3d_pkg
import make_mapmerging_golden as gen
mm = mm3d_pkg.load()
ctx = mm.Context(0)
for name, ds, dt in gen.match_cases():
    for k in (1, 5, 8):
        for a, b, tag in ((ds, dt, "fwd"), (dt, ds, "bwd")):
            kk = min(k, len(b))
            os.environ["MM3D_KNN"] = "exact"; ie, de = ctx.knn(a, b, kk)
            os.environ["MM3D_KNN"] = "tc"; it, dt_ = ctx.knn(a, b, kk)
            bad = np.where((ie != it).any(axis=1) | (de.view(np.uint32) != dt_.view(np.uint32)).any(axis=1))[0]
            print(f"{name} k={k} {tag}: {len(a)} x {len(b)} dim {a.shape[1]}: {len(bad)} rows differ")
            for r in bad[:3]:
                print(f"   row {r}: exact {ie[r].tolist()} {de[r].tolist()}\n           tc    {it[r].tolist()} {dt_[r].tolist()}")
