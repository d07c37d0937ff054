#!/usr/bin/env python
"""Extracts per-launch DRAM traffic and headline counters of the profiled kernels from an
.ncu-rep (read here, no GPU) and merges them into profiles/traffic.json, which bench.py uses
for roofline.traffic:

    python tools/ncu_traffic.py gpurun_out/prof_c2_v4.ncu-rep --workload c2 --rays 4190209 --labels primary,bounce
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
    ap.add_argument("--rays", type=int, required=True, help="rays per profiled launch")
    ap.add_argument("--labels", required=True, help="comma separated names of the profiled launches, in order")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "traffic.json"))
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(hdr)}
    data = json.load(open(args.out)) if os.path.exists(args.out) else {}
    entry = {"source": os.path.basename(args.report), "rays_per_profiled_launch": args.rays, "launches": {}}
    for label, r in zip(args.labels.split(","), rows[2:]):
        def val(k):
            return float(r[col[k]].replace(",", "")) * UNITS.get(units[col[k]], 1.0)
        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        launch = {"kernel": r[col["Kernel Name"]], "dram_read_bytes": rd, "dram_write_bytes": wr,
                  "dram_bytes_per_ray": (rd + wr) / args.rays}
        for k in KEYS:
            if k in col:
                launch[k] = float(r[col[k]].replace(",", ""))
        entry["launches"][label] = launch
    data[args.workload] = entry
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(data, open(args.out, "w"), indent=1, sort_keys=True)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
