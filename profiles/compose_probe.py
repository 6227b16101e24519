"""Wall time of consecutive composeMaps calls on the c5 workload (resident and host-buffer paths), call by call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mm3d_pkg
mm = mm3d_pkg.load(); synth = mm3d_pkg.load_synth()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg = dict(synth.CONFIGS["c5"]); cfg["n_maps"] = n
maps, _ = synth.make_maps(**cfg)
ctx = mm.Context(0)
T = np.stack([np.eye(4, dtype=np.float32)] * n)
dm = ctx.maps_upload(maps)
for name, fn in (("resident", lambda: ctx.compose_resident_dist(None, dm, T, 0.05)), ("host", lambda: ctx.compose_maps_dist(None, maps, T, 0.05)),
                 ("resident", lambda: ctx.compose_resident_dist(None, dm, T, 0.05)), ("plain", lambda: ctx.compose_maps(maps, T, 0.05))):
    ts = []
    for _ in range(5):
        t0 = time.time(); out = fn(); ts.append((time.time() - t0) * 1e3)
    print(name, " ".join(f"{t:.1f}" for t in ts), "ms; output", out.shape, flush=True)
