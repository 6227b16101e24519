"""One resident step of a workload (for ncu, or for a one-off timing of the bigger BASELINE configs):
1 warm-up step + N timed steps; prints sizes, per-stage device milliseconds and wall time."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mm3d_pkg

mm = mm3d_pkg.load(); synth = mm3d_pkg.load_synth()
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
t0 = time.time()
cfg = dict(synth.CONFIGS[name])
if os.environ.get("MM3D_NMAPS"):  # the same map size with fewer maps (short ncu captures)
    cfg["n_maps"] = int(os.environ["MM3D_NMAPS"])
maps, _ = synth.make_maps(**cfg)
print(f"generated {len(maps)} maps x {len(maps[0])} points in {time.time() - t0:.1f} s", flush=True)
ctx = mm.Context(0)
p = mm.default_params(descriptor_type="FPFH")
dm = ctx.maps_upload(maps)
l0 = ctx.launches
ctx.estimate_resident(dm, p)
print("launches per step:", ctx.launches - l0, flush=True)
for _ in range(steps):
    t0 = time.time()
    T, st = ctx.estimate_resident(dm, p, stage_times=True)
    dt = time.time() - t0
    n = len(maps)
    print(f"step: {dt * 1e3:.1f} ms wall, {n * (n - 1) // 2 / dt:.1f} pairs/s, transforms {T.shape}; stage ms: " +
          ", ".join(f"{k}={v:.1f}" for k, v in st.items()), flush=True)
f = ctx.features_compute(dm, 0, min(len(maps), 4), p)
npt, nk, dim = f.sizes()
print("points after downSample+removeOutliers:", npt.tolist(), "keypoints:", nk.tolist(), "dim:", dim)
