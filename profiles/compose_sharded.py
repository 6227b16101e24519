"""composeMaps sharded over the ranks of one node (BASELINE config 5 shape): every rank holds
M / world maps of P points, the ranks exchange raw points by voxel-key range and voxel-grid their own range.
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 profiles/compose_sharded.py [M P res steps]
Prints per-step device-side wall (max over ranks), points in / out, and (N = 1 .. small sizes) a check against the
unsharded library call."""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import mm3d_pkg

mm = mm3d_pkg.load()
sh = importlib.import_module("map_merge_b200.sharding")
synth = importlib.import_module("map_merge_b200.synth")

M = int(sys.argv[1]) if len(sys.argv) > 1 else 8
P = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
res = float(sys.argv[3]) if len(sys.argv) > 3 else 0.05
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ctx = mm.Context(local)
first, count, _ = sh.map_block(rank, world, M)
all_maps, all_truth = synth.make_maps(11, M, P, 60.0, 30.0, 6, 3)
maps = all_maps[first:first + count]
T_all = np.stack([np.linalg.inv(all_truth[0]) @ t for t in all_truth]).astype(np.float32)
T = T_all[first:first + count]
check = os.environ.get("CHECK") == "1"
if not (check and rank == 0):
    del all_maps
times = []
for s in range(steps + 1):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = sh.compose_sharded(sh.CtxShardOps(ctx, torch, dev), dist if world > 1 else None, torch, dev, maps, T, res)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if s > 0:
        times.append(float(dt.item()))
n_out = torch.tensor([len(out)], device=dev)
if world > 1:
    dist.all_reduce(n_out)
if rank == 0:
    ms = 1e3 * float(np.mean(times))
    print(f"compose_sharded: {M} maps x {P} points, resolution {res}, {world} rank(s): {ms:.1f} ms/step, "
          f"{M * P / ms / 1e3:.1f} Mpoints/s in, {int(n_out.item())} points out (this rank {len(out)})")
if check:
    # gather every rank's slice on rank 0 and compare with the unsharded library call, bit for bit
    sizes = torch.zeros(world, dtype=torch.int64, device=dev); sizes[rank] = len(out)
    if world > 1:
        dist.all_reduce(sizes)
    buf = torch.zeros((int(sizes.max().item()), 4), dtype=torch.float32, device=dev)
    buf[:len(out)] = torch.from_numpy(out).to(dev)
    parts = [torch.zeros_like(buf) for _ in range(world)]
    if world > 1:
        dist.all_gather(parts, buf)
    else:
        parts = [buf]
    if rank == 0:
        got = np.concatenate([parts[r][:int(sizes[r].item())].cpu().numpy() for r in range(world)])
        want = ctx.compose_maps(all_maps, T_all, res)
        same = got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))
        print("sharded == unsharded (bit-exact):", same)
        assert same
if world > 1:
    dist.destroy_process_group()
