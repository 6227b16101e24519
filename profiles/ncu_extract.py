"""Summarise .ncu-rep captures (gpurun_out/) into profiles/: one CSV row per captured kernel launch with the metrics the
design discussion cites, and a traffic JSON (DRAM bytes per launch) for bench.py's roofline.traffic.  Runs on the CPU box
(`ncu -i`).  Usage: python profiles/ncu_extract.py out_prefix rep1.ncu-rep [rep2 ...]"""
import csv, json, subprocess, sys, os

METRICS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def to_bytes(v, unit):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    prefix, reps = sys.argv[1], sys.argv[2:]
    rows_out, traffic = [], {}
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")]
            short = name.split("(")[0].split("::")[-1]
            if "<" in name:
                short = name[name.index(short):].split("(const")[0].strip()
            r = {"report": os.path.basename(rep), "kernel": short}
            for m in METRICS:
                if m in hdr:
                    i = hdr.index(m)
                    r[f"{m} [{units[i]}]"] = vals[i]
            rows_out.append(r)
            try:
                rd = to_bytes(vals[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
                wr = to_bytes(vals[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
                traffic.setdefault(short, []).append(rd + wr)
            except ValueError:
                pass
    keys = []
    for r in rows_out:
        for k in r:
            if k not in keys:
                keys.append(k)
    with open(prefix + "_ncu_summary.csv", "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=keys)
        w.writeheader()
        w.writerows(rows_out)
    print("wrote", prefix + "_ncu_summary.csv", len(rows_out), "rows")
    for k, v in traffic.items():
        print(k, [f"{x / 1e6:.1f} MB" for x in v])


if __name__ == "__main__":
    main()
