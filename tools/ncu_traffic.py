#!/usr/bin/env python
"""Extracts per-launch DRAM traffic and headline counters of the profiled kernels from an
.ncu-rep (read here, no GPU) and merges them into profiles/traffic.json, which bench.py uses
for roofline.traffic:

    python tools/ncu_traffic.py gpurun_out/prof_c2_v4.ncu-rep --workload c2 --rays 4190209 --labels primary,bounce

Several launches per label (the path stream: one closest-hit and one probe launch per depth), rays
summed over them:

    python tools/ncu_traffic.py gpurun_out/prof_c5.ncu-rep --workload c5 --aggregate closest:0,2,4,6:RAYS probe:1,3,5,7:RAYS

The entry records the hash of the kernel sources the capture was taken from (bench.kernel_source_sha):
bench.py uses a capture only for the kernels it was taken from.
"""
import argparse
import csv
import io
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
KEYS = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--workload", required=True)
    ap.add_argument("--rays", type=int, default=0, help="rays per profiled launch")
    ap.add_argument("--labels", default="", help="comma separated names of the profiled launches, in order")
    ap.add_argument("--aggregate", nargs="*", default=[], help="label:launch,launch,...:rays -- several profiled launches summed into one label")
    ap.add_argument("--kernel-sha", default=None, help="hash of the kernel sources of the capture (default: the current sources)")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "traffic.json"))
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(hdr)}
    data = json.load(open(args.out)) if os.path.exists(args.out) else {}
    import sys
    sys.path.insert(0, ROOT)
    import bench
    entry = {"source": os.path.basename(args.report), "rays_per_profiled_launch": args.rays, "launches": {},
             "kernel_source_sha": args.kernel_sha or bench.kernel_source_sha()}

    def val(r, k):
        return float(r[col[k]].replace(",", "")) * UNITS.get(units[col[k]], 1.0)

    groups = [(label, [i], args.rays) for i, label in enumerate(args.labels.split(",")) if label]
    for spec in args.aggregate:
        label, which, rays = spec.split(":")
        groups.append((label, [int(x) for x in which.split(",")], int(rays)))
    for label, which, rays in groups:
        rs = [rows[2 + i] for i in which]
        rd, wr = sum(val(r, "dram__bytes_read.sum") for r in rs), sum(val(r, "dram__bytes_write.sum") for r in rs)
        time_ms = sum(val(r, "gpu__time_duration.sum") for r in rs)
        launch = {"kernel": rs[0][col["Kernel Name"]], "profiled_launches": len(rs), "rays": rays, "dram_read_bytes": rd, "dram_write_bytes": wr,
                  "dram_bytes_per_ray": (rd + wr) / rays}
        for k in KEYS:
            if k in col:
                # duration and instruction count add up; the rest is averaged, weighted by duration
                if k in ("gpu__time_duration.sum", "smsp__inst_executed.sum"):
                    launch[k] = sum(float(r[col[k]].replace(",", "")) for r in rs)
                else:
                    launch[k] = sum(float(r[col[k]].replace(",", "")) * val(r, "gpu__time_duration.sum") for r in rs) / max(time_ms, 1e-30)
        entry["launches"][label] = launch
    data[args.workload] = entry
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(data, open(args.out, "w"), indent=1, sort_keys=True)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
