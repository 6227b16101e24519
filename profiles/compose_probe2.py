"""Where does a composeMaps step of config 5 spend its time?  Wall clock around the pieces of the bench step."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mm3d_pkg
mm = mm3d_pkg.load(); synth = mm3d_pkg.load_synth()
cfg = dict(synth.CONFIGS["c5"])
maps, _ = synth.make_maps(**cfg)
ctx = mm.Context(0)
res = ctx.maps_upload(maps)
T = np.stack([np.eye(4, dtype=np.float32)] * len(maps))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda:0")
def sync(): torch.cuda.synchronize()
for it in range(6):
    sync(); t0 = time.perf_counter()
    flush.zero_(); sync(); t1 = time.perf_counter()
    out = ctx.compose_resident_dist(None, res, T, 0.05); t2 = time.perf_counter()
    sync(); t3 = time.perf_counter()
    del out; t4 = time.perf_counter()
    print(f"iter {it}: flush {1e3*(t1-t0):.2f} ms, compose call {1e3*(t2-t1):.2f} ms, trailing sync {1e3*(t3-t2):.2f} ms, release {1e3*(t4-t3):.2f} ms", flush=True)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
sync(); ev0.record()
for _ in range(5):
    flush.zero_(); out = ctx.compose_resident_dist(None, res, T, 0.05)
ev1.record(); sync()
print("event-timed loop:", ev0.elapsed_time(ev1) / 5, "ms per step")
sync(); ev0.record()
for _ in range(5):
    out = ctx.compose_resident_dist(None, res, T, 0.05)
ev1.record(); sync()
print("event-timed loop without the flush:", ev0.elapsed_time(ev1) / 5, "ms per step")
