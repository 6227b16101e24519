"""One resident step of the bench workload (for ncu): 1 warm-up step + N profiled steps."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mm3d_pkg

mm = mm3d_pkg.load(); synth = mm3d_pkg.load_synth()
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
maps, _ = synth.make_maps(**synth.CONFIGS[name])
ctx = mm.Context(0)
p = mm.default_params(descriptor_type="FPFH")
dm = ctx.maps_upload(maps)
l0 = ctx.launches
ctx.estimate_resident(dm, p)
print("launches per step:", ctx.launches - l0)
for _ in range(steps):
    ctx.estimate_resident(dm, p)
