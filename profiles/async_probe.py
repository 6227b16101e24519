"""Why is a c3 step slower without stage synchronisation?  One resident c3 step in three modes, wall time per step,
with MM3D_HOST_TRACE=1 printing where the host spends its time."""
import os, sys, time
os.environ["MM3D_HOST_TRACE"] = "1"  # read once by the library: per-stage host time + time inside the runtime calls
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mm3d_pkg
mm = mm3d_pkg.load(); synth = mm3d_pkg.load_synth()
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
maps, _ = synth.make_maps(**synth.CONFIGS[name])
ctx = mm.Context(0)
p = mm.default_params(descriptor_type="FPFH")
dm = ctx.maps_upload(maps)
for _ in range(3):
    ctx.estimate_resident(dm, p)
for mode in ("plain", "stage_times", "plain", "profile", "plain"):
    if mode == "profile":
        ctx.profile_begin()
    t0 = time.time()
    for _ in range(2):
        ctx.estimate_resident(dm, p, stage_times=(mode == "stage_times"))
    dt = (time.time() - t0) / 2
    if mode == "profile":
        ctx.profile_end()
    print(f"{mode}: {dt * 1e3:.1f} ms per step", flush=True)
for _ in range(4):
    t0 = time.time()
    ctx.estimate_resident(dm, p)
    print(f"traced step: {(time.time() - t0) * 1e3:.1f} ms", flush=True)
