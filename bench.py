#!/usr/bin/env python
"""bench.py — pairwise registrations/sec of the map_merge_3d hot path on B200 (and the CPU reference arm).

A "step" is one whole estimateMapsTransforms() over the workload: the per-map feature pipeline
(voxel grid, outlier removal, normals, SIFT3D, FPFH) for every map plus the all-pairs loop
(reciprocal k-NN matching, RANSAC, ICP refine, scoring) and the host pose graph.
Workload = BASELINE.json configs[1]: 8 synthetic overlapping maps, 500k points each, FPFH, 28 pairs.

  value : pairs/s, inputs already resident in HBM when the timed region starts (CUDA events, max over ranks)
  e2e   : pairs/s through the C-ABI call with HOST buffers (pinned), H2D of the clouds and D2H of the
          transforms inside the timed region
  N > 1 : one process per GPU (torchrun); maps are sharded for the feature pipeline, features are exchanged
          with NCCL all_gather, the pair list is sharded, pair results are all_gathered, rank 0 runs the
          host graph.  Same workload at every N => "scaling": "strong".
  --impl reference : the CPU restatement of the reference (oracle/) on all host threads, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "pairwise_registrations_per_sec"
UNIT = "pairs/s"


def workload(name: str):
    import mm3d_pkg
    synth = mm3d_pkg.load_synth()
    cfg = dict(synth.CONFIGS[name])
    maps, truth = synth.make_maps(**cfg)
    return maps, truth, cfg


def n_pairs_of(m: int) -> int:
    return m * (m - 1) // 2


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------- CPU arms
def oracle_parallel_job(O, maps, p, threads: int):
    """The reference's path (CPU restatement) with the map loop and the pair loop on a thread pool."""
    from concurrent.futures import ThreadPoolExecutor

    def feat(m):
        t0 = time.perf_counter()
        ds, _ = O.downsample(m, p.resolution)
        fo, _, _ = O.remove_outliers(ds, p.descriptor_radius, p.outliers_min_neighbours)
        nm = O.normals(fo, p.normal_radius)
        kp = O.sift(fo, p.resolution, p.keypoint_threshold)
        kp, desc = O.fpfh(fo, nm, kp, p.descriptor_radius)
        return dict(cloud=fo, kp=kp, desc=desc, t=time.perf_counter() - t0)

    def pair(a, b):
        t0 = time.perf_counter()
        pr, dist = O.match(a["desc"], b["desc"], p.matching_k)
        T, inl, _ = O.ransac(a["kp"], b["kp"], pr, dist, p.inlier_threshold)
        if p.refine_transform:
            T, _ = O.icp(a["cloud"], b["cloud"], T, p.max_correspondence_distance, p.max_iterations, p.transform_epsilon)
        s = O.score(a["cloud"], b["cloud"], T, p.max_correspondence_distance)
        return T, 1.0 / s, time.perf_counter() - t0

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        feats = list(ex.map(feat, maps))
        t1 = time.perf_counter()
        ij = [(i, j) for i in range(len(maps) - 1) for j in range(i + 1, len(maps)) if len(feats[i]["kp"]) and len(feats[j]["kp"])]
        res = list(ex.map(lambda t: pair(feats[t[0]], feats[t[1]]), ij))
    t2 = time.perf_counter()
    if ij:
        O.global_transforms(np.array(ij, np.int32), np.stack([r[0] for r in res]), [r[1] for r in res], p.confidence_threshold)
    return dict(wall=time.perf_counter() - t0, feat_wall=t1 - t0, pair_wall=t2 - t1, feat_t=[f["t"] for f in feats],
                pair_t=[r[2] for r in res], n_pairs=len(ij))


def extrapolate(feat_t, pair_t, n_maps, n_pairs, threads):
    """Full-job time from per-map / per-pair times measured on a sample, for `threads` workers."""
    f = float(np.mean(feat_t)); g = float(np.mean(pair_t)) if len(pair_t) else 0.0
    return -(-n_maps // threads) * f + -(-n_pairs // threads) * g


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    O = oracle_py.Oracle()
    maps, truth, cfg = workload(args.workload)
    p = oracle_py.default_params(descriptor_type=2)
    cores = os.cpu_count() or 1
    M = len(maps)
    P = n_pairs_of(M)
    # bounded sample: as many maps per step as fit ~200 s for the whole run (12.5 s/map, 1.2 s/pair measured at 500k pts)
    total_steps = args.steps + args.warmup
    budget = 200.0 / max(total_steps, 1)
    sample = 2
    for m in (8, 4, 3, 2):
        if m > M:
            continue
        est = -(-m // cores) * 13.0 + -(-n_pairs_of(m) // cores) * 1.5
        if est <= budget:
            sample = m
            break
    sub = maps[:sample]
    times, last = [], None
    for s in range(total_steps):
        r = oracle_parallel_job(O, sub, p, cores)
        if s >= args.warmup:
            times.append(r)
        last = r
    full = float(np.mean([extrapolate(r["feat_t"], r["pair_t"], M, P, cores) if sample < M else r["wall"] for r in times]))
    value = P / full
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": full * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {M} maps x {cfg['n_points']} pts, SIFT+FPFH, MATCHING+ICP, {P} pairs", "seed": cfg["seed"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": (f"maps 0-{sample - 1} of the workload ({n_pairs_of(sample)} of {P} pairs) per step, map loop and pair loop on "
                                    f"{cores} threads; per-map and per-pair times extrapolated to {M} maps / {P} pairs" if sample < M else
                                    f"whole workload on {cores} threads")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _OUT.write(json.dumps(line) + "\n")
    _OUT.flush()


def cpu_baseline_sample(maps, M, P):
    """Single-thread oracle on maps 0-1 of the workload (one pair), extrapolated to the full job."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py
    O = oracle_py.Oracle()
    p = oracle_py.default_params(descriptor_type=2)
    t0 = time.perf_counter()
    r = O.estimate_maps_transforms(maps[:2], p)
    wall = time.perf_counter() - t0
    st = r["stage_times"]
    feat = sum(st[k] for k in ("downsampling", "removing outliers", "normals computation", "keypoints detection", "descriptors computation")) / 2
    pair = sum(st[k] for k in ("finding correspondences", "initial alignment", "ICP alignment", "scoring"))
    full = M * feat + P * pair
    return {"value": P / full, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"maps 0-1 of the workload (1 of {P} pairs), {wall:.1f} s single-thread; per-map ({feat:.2f} s) and per-pair ({pair:.2f} s) "
                      f"stage times extrapolated to {M} maps / {P} pairs",
            "stage_seconds_sample": {k: round(v, 4) for k, v in st.items()}}, r


# ---------------------------------------------------------------------------------------------- GPU arm
class Job:
    """One rank's share of the sharded path."""

    def __init__(self, args, maps):
        import torch
        import torch.distributed as dist
        import mm3d_pkg
        self.torch, self.dist = torch, dist
        self.mm = mm3d_pkg.load()
        import importlib
        self.sh = importlib.import_module("map_merge_b200.sharding")
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.phases_on = os.environ.get("MM3D_BENCH_PHASES") == "1"
        self.phase_ms = {}
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.ctx = self.mm.Context(self.local, stream=torch.cuda.current_stream().cuda_stream)
        self.p = self.mm.default_params(descriptor_type="FPFH")
        self.M = len(maps)
        self.P = n_pairs_of(self.M)
        # contiguous block of maps per rank
        self.first, self.count, per = self.sh.map_block(self.rank, self.world, self.M)
        self.owner_of_map = [self.sh.owner_of_map(m, self.world, self.M) for m in range(self.M)]
        # pinned host copies of this rank's maps (e2e path) and resident device copies (value path)
        self.host = []
        for m in range(self.first, self.first + self.count):
            t = torch.empty((len(maps[m]), 4), dtype=torch.float32, pin_memory=True)
            t.numpy()[:] = maps[m]
            self.host.append(t)
        self.resident = self.ctx.maps_upload([t.numpy() for t in self.host])
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)
        self.h2d_bytes = sum(int(t.numel()) * 4 for t in self.host)

    # -- single GPU --------------------------------------------------------
    def step_single(self, from_host: bool):
        if from_host:
            return self.ctx.estimate_maps_transforms([t.numpy() for t in self.host], self.p)
        return self.ctx.estimate_resident(self.resident, self.p)

    # -- sharded -----------------------------------------------------------
    def _phase(self, name):
        """MM3D_BENCH_PHASES=1: synchronising per-phase wall times on every rank (diagnostic; perturbs the step)."""
        if not self.phases_on:
            return
        self.torch.cuda.synchronize()
        now = time.perf_counter()
        self.phase_ms[name] = self.phase_ms.get(name, 0.0) + (now - self._t_phase) * 1e3
        self._t_phase = now

    def step_sharded(self, from_host: bool):
        torch, dist = self.torch, self.dist
        if self.phases_on:
            torch.cuda.synchronize(); self._t_phase = time.perf_counter()
        maps = self.ctx.maps_upload([t.numpy() for t in self.host]) if from_host else self.resident
        feats = self.ctx.features_compute(maps, 0, self.count, self.p)
        self._phase("features")
        npt, nk, dim = feats.sizes()
        per = -(-self.M // self.world)
        sizes = torch.zeros((per, 2), dtype=torch.int32, device=self.dev)
        if self.count:
            sizes[:self.count, 0] = torch.from_numpy(npt).to(self.dev)
            sizes[:self.count, 1] = torch.from_numpy(nk).to(self.dev)
        all_sizes = torch.empty((self.world, per, 2), dtype=torch.int32, device=self.dev)
        dist.all_gather_into_tensor(all_sizes, sizes)
        all_sizes = all_sizes.cpu().numpy()
        max_pt = int(all_sizes[:, :, 0].max()); max_kp = int(all_sizes[:, :, 1].max())
        # one padded buffer per rank: [per][points | keypoints | descriptors]
        row = max_pt * 4 + max_kp * 4 + max_kp * dim
        send = torch.zeros((per, row), dtype=torch.float32, device=self.dev)
        for m in range(self.count):
            base = send[m].data_ptr()
            feats.export_dev(m, base, base + max_pt * 16, base + (max_pt + max_kp) * 16)
        recv = torch.empty((self.world, per, row), dtype=torch.float32, device=self.dev)
        dist.all_gather_into_tensor(recv, send)
        self._phase("exchange")
        n_points, n_kp, pp, kp, dp = [], [], [], [], []
        for m in range(self.M):
            r, l = self.owner_of_map[m], m - self.owner_of_map[m] * per
            base = recv[r, l].data_ptr()
            n_points.append(int(all_sizes[r, l, 0])); n_kp.append(int(all_sizes[r, l, 1]))
            pp.append(base); kp.append(base + max_pt * 16); dp.append(base + (max_pt + max_kp) * 16)
        torch.cuda.current_stream().synchronize()
        allf = self.ctx.features_import_dev(n_points, pp, n_kp, kp, dp, dim)
        ij = self.sh.pair_list(n_kp)
        owner = self.sh.lpt_assign(self.sh.pair_costs(ij, n_points, n_kp, dim), self.world) if ij else np.zeros(0, np.int64)
        mine = [k for k in range(len(ij)) if owner[k] == self.rank]
        self._phase("import")
        T, conf, stats = self.ctx.register_pairs(allf, [ij[k] for k in mine], self.p)
        self._phase("register")
        # every rank fills its own slots of the row-major pair list (keeps the reference's pair order)
        res = self.sh.gather_pair_results(dist, torch, self.dev, len(ij), mine, T, conf)
        out = None
        if self.rank == 0 and len(ij):
            h = res.cpu().numpy()
            out, _ = self.mm.global_transforms(np.array(ij, np.int32), h[:, :16].reshape(-1, 4, 4).astype(np.float32), h[:, 16],
                                               self.p.confidence_threshold)
        self._phase("results+graph")
        return out

    def step(self, from_host: bool):
        return self.step_single(from_host) if self.world == 1 else self.step_sharded(from_host)

    def timed(self, steps: int, from_host: bool, profile: bool):
        torch, dist = self.torch, self.dist
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = self.ctx.launches
        if profile:
            self.ctx.profile_begin()
        ev0.record()
        out = None
        for _ in range(steps):
            self.flush.zero_()  # evict the previous step's working set from L2
            out = self.step(from_host)
        ev1.record()
        torch.cuda.synchronize()
        prof = self.ctx.profile_end() if profile else None
        if self.world > 1:
            dist.barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), self.ctx.launches - l0, prof, out


def run_gpu(args):
    maps, truth, cfg = workload(args.workload)
    job = Job(args, maps)
    M, P = job.M, job.P
    for _ in range(max(args.warmup, 3)):
        job.step(False)
    sampler = ClockSampler(job.local)
    if job.rank == 0:
        sampler.start()  # nvidia-smi needs ~0.5 s to produce its first sample: start it one (identical, untimed) pass early
    job.timed(max(args.steps, 10), from_host=False, profile=True)  # untimed: also fills the library's CUDA-event pool
    job.phase_ms = {}
    ms, launches, prof, out = job.timed(args.steps, from_host=False, profile=True)
    if job.phases_on:
        print(f"[phases] rank {job.rank}: " + ", ".join(f"{k} {v / args.steps:.2f} ms" for k, v in job.phase_ms.items()) +
              f"; step {ms / args.steps:.2f} ms", file=sys.stderr, flush=True)
    clocks = sampler.stop() if job.rank == 0 else None
    job.step(True)  # warm the host path
    ms_e2e, _, _, out_e2e = job.timed(args.steps, from_host=True, profile=False)
    if job.rank != 0:
        return
    value = P * args.steps / (ms * 1e-3)
    e2e = P * args.steps / (ms_e2e * 1e-3)
    peak, peak_src = measured_peak()
    # dominant kernel by device time over the timed region
    prof = sorted(prof, key=lambda k: -k["ms"])
    total_kernel_ms = sum(k["ms"] for k in prof)
    top = prof[0]
    roof = {"bound": "hbm", "kernel": top["kernel"], "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None,
            "peak_source": peak_src, "launches": top["launches"], "avg_launch_ms": top["ms"] / max(top["launches"], 1),
            "share_of_kernel_time": top["ms"] / max(total_kernel_ms, 1e-9),
            "note": "achieved = algorithmic bytes per launch (DESIGN.md, SURVEY.md 8d) / CUDA-event launch time; neighbourhood kernels are "
                    "L2-gather bound, so their HBM fraction is small by construction"}
    if top["algorithmic_bytes"] > 0 and top["ms_annotated"] > 0:
        roof["achieved"] = top["algorithmic_bytes"] / (top["ms_annotated"] * 1e-3) / 1e9
        roof["frac"] = roof["achieved"] / peak
        roof["algorithmic_bytes_per_launch"] = top["algorithmic_bytes"] / top["launches"]
    # what actually bounds the kernel (ncu --set full capture of the same kernel, committed under profiles/)
    try:
        import csv
        with open(os.path.join(ROOT, "profiles", "r01_ncu_summary_v2.csv")) as fh:
            rows = [r for r in csv.DictReader(fh) if r["kernel"] == top["kernel"]]
        if rows:
            r0 = max(rows, key=lambda r: float(r["gpu__time_duration.sum [ms]"]))
            roof["ncu"] = {"issue_slots_busy_pct": float(r0["smsp__issue_active.avg.pct_of_peak_sustained_active [%]"]),
                           "active_lanes_per_warp": float(r0["smsp__thread_inst_executed_per_inst_executed.ratio []"]),
                           "l1_hit_pct": float(r0["l1tex__t_sector_hit_rate.pct [%]"]),
                           "source": "profiles/r01_ncu_summary_v2.csv (issue bound, not HBM bound)"}
    except Exception:
        pass
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            roof["traffic"] = json.load(open(traffic_file)).get(top["kernel"])
        except Exception:
            pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": job.world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {M} maps x {cfg['n_points']} pts, SIFT+FPFH, MATCHING+ICP, {P} pairs", "seed": cfg["seed"],
                   "l2": f"256 MiB memset between steps (inputs are {M * cfg['n_points'] * 16 // 1000000} MB)",
                   "parallelism": "single GPU" if job.world == 1 else f"maps and pairs sharded over {job.world} ranks, NCCL all_gather of features"},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(job.h2d_bytes), "d2h_bytes_per_step": int(M * 64),
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "kernels": [{"kernel": k["kernel"], "launches": k["launches"], "ms_per_step": k["ms"] / args.steps,
                     "gbps": (k["algorithmic_bytes"] / (k["ms_annotated"] * 1e-3) / 1e9) if k["ms_annotated"] > 0 else None} for k in prof[:12]],
    }
    if job.world == 1 and not args.no_cpu_baseline:
        cb, ref = cpu_baseline_sample(maps, M, P)
        line["cpu_baseline"] = cb
    else:
        line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "timed at N=1 only"}
    _OUT.write(json.dumps(line) + "\n")
    _OUT.flush()
    if job.world > 1:
        job.dist.destroy_process_group()


_OUT = sys.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mm3d", choices=["mm3d", "reference"])
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: keep a private handle to it and point fd 1 at stderr, so that anything a
    # library writes to stdout (NCCL prints its version banner there) cannot get in front of the line
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
