"""Where do the tensor-core filter's candidates sit in the distance matrix?  Per (32-row block, 32-column chunk) cell: how
many candidates, and which share of all candidates lies in dense cells.  GPU for the descriptors, numpy (float64) after."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mm3d_pkg
mm = mm3d_pkg.load()
synth = importlib.import_module("map_merge_b200.synth")
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
ma, mb = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 4)
maps, _ = synth.make_maps(**synth.CONFIGS[name], only=[ma, mb])
ctx = mm.Context(0)
p = mm.default_params(descriptor_type="FPFH")
f = ctx.features_compute(ctx.maps_upload([maps[ma], maps[mb]]), 0, 2, p)
a, b = [np.asarray(f.export_host(i)[2], np.float64) for i in range(2)]
mu = np.concatenate([a, b]).mean(0)
A, B = a - mu, b - mu
na, nb = (A * A).sum(1), (B * B).sum(1)
D = na[:, None] + nb[None, :] - 2 * A @ B.T
kth = np.partition(D, 4, axis=1)[:, 4]
ES = 3.1e-5
C = (D - ES * (na[:, None] + nb[None, :])) <= kth[:, None]
per_row = C.sum(1)
print(f"{name} maps {ma},{mb}: {C.shape}; candidates/row mean {per_row.mean():.1f} median {np.median(per_row):.0f} p90 {np.percentile(per_row, 90):.0f} max {per_row.max()}")
for E in (3.1e-5, 1e-5, 3e-6, 0.0):
    pr = ((D - E * (na[:, None] + nb[None, :])) <= kth[:, None]).sum(1)
    print(f"  ES={E:g}: candidates/row mean {pr.mean():.1f}")
R, Cc = (C.shape[0] + 31) // 32, (C.shape[1] + 31) // 32
P = np.zeros((R * 32, Cc * 32), bool); P[:C.shape[0], :C.shape[1]] = C
cells = P.reshape(R, 32, Cc, 32).sum(axis=(1, 3))
tot = cells.sum()
print(f"cells {cells.size}, non-empty {np.count_nonzero(cells)} ({np.count_nonzero(cells) / cells.size * 100:.1f}%)")
for t in (32, 64, 128, 256, 512, 768):
    m = cells >= t
    print(f"  cells with >= {t:4d} of 1024 candidates: {m.sum():6d} ({m.mean() * 100:5.2f}% of cells) holding {cells[m].sum() / tot * 100:5.1f}% of the candidates; dense work / sparse work there = {m.sum() * 1024 / max(cells[m].sum(), 1):.2f}")
# per row: rows sorted by candidate count
srt = np.sort(per_row)[::-1]
cum = np.cumsum(srt) / srt.sum()
for q in (0.01, 0.05, 0.1, 0.2, 0.5):
    print(f"  top {q * 100:.0f}% of rows hold {cum[int(q * len(srt)) - 1] * 100:.1f}% of the candidates")
# lane = row alternative: per block of 32 consecutive rows, the union of the rows' candidate columns
U = P.reshape(R, 32, -1).any(axis=1).sum(axis=1)  # union size per row block
S = P.reshape(R, 32, -1).sum(axis=(1, 2))          # candidates per row block
print(f"row blocks {R}: union columns per block mean {U.mean():.0f} median {np.median(U):.0f} p90 {np.percentile(U, 90):.0f} max {U.max()};"
      f" total union columns {U.sum()} vs candidates {S.sum()} ({S.sum() / 32 / U.sum() * 100:.1f}% lane efficiency)")
heavy = S > 32 * 100
print(f"  heavy blocks (> 100 candidates/row): {heavy.sum()} with union mean {U[heavy].mean():.0f}, candidates/row {S[heavy].mean() / 32:.0f};"
      f" light blocks: {(~heavy).sum()} with union mean {U[~heavy].mean():.0f}, candidates/row {S[~heavy].mean() / 32:.1f}")
for rb in (8, 16):
    Rb = (C.shape[0] + rb - 1) // rb
    Pb = np.zeros((Rb * rb, P.shape[1]), bool); Pb[:C.shape[0], :C.shape[1]] = C
    Ub = Pb.reshape(Rb, rb, -1).any(axis=1).sum(axis=1)
    print(f"  blocks of {rb} rows: total union columns {Ub.sum()} ({C.sum() / rb / Ub.sum() * 100:.1f}% efficiency)")
