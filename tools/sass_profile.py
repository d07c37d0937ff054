#!/usr/bin/env python
"""Per-source-line view of an ncu capture (no GPU needed).

Joins the SASS page of an .ncu-rep (per-instruction executed counts and stall samples) with
`nvdisasm --print-line-info` of the cubin the library carries, by instruction order, and prints
the source lines ranked by executed warp instructions / stall samples.

    python tools/sass_profile.py gpurun_out/prof_c2.ncu-rep --block 1 \\
        --func 'trace_kernelILb0ELb1ELb0E' [--lib appleseed_b200/libasgpu.so] [--top 40]
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import tempfile


def disasm_lines(lib, func, outer=False):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    out = []
    for name in sorted(os.listdir(tmp)):
        if not name.endswith(".cubin"):
            continue
        text = subprocess.run(["nvdisasm", "--print-line-info-inline" if outer else "--print-line-info", os.path.join(tmp, name)], capture_output=True, text=True).stdout
        cur, inside, line = None, False, ("?", 0)
        for l in text.splitlines():
            if l.startswith(".text."):
                inside = func in l
                cur = []
                if inside:
                    out.append(cur)
                continue
            if not inside:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                # With inline info the chain runs innermost first; the last entry before the
                # instruction is the outermost call site (a line of the kernel itself).
                line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
            if m:
                cur.append((int(m.group(1), 16), m.group(2).strip(), line))
    return max(out, key=len) if out else []


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--block", type=int, default=0, help="n-th profiled launch in the report (0-based)")
    ap.add_argument("--func", required=True, help="substring of the mangled kernel name")
    ap.add_argument("--lib", default="appleseed_b200/libasgpu.so")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--sass", action="store_true", help="also list the hottest SASS instructions")
    ap.add_argument("--outer", action="store_true", help="attribute inlined code to the outermost call site (a line of the kernel)")
    ap.add_argument("--regions", default="", help="with --outer: 'name=first-last,...' kernel line ranges summed into named regions")
    args = ap.parse_args()

    raw = subprocess.run(["ncu", "-i", args.report, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    blocks = []
    for k, hi in enumerate(starts):
        end = starts[k + 1] - 1 if k + 1 < len(starts) else len(rows)
        blk = (rows[hi - 1][1], rows[hi], [r for r in rows[hi + 1:end] if len(r) == len(rows[hi])])
        if not blocks or blk[2] != blocks[-1][2]:       # ncu prints every launch twice
            blocks.append(blk)
    name, hdr, body = blocks[args.block]
    print("launch %d of %d: %s" % (args.block, len(blocks), name[:100]))
    col = {n: hdr.index(n) for n in hdr}
    sass = disasm_lines(args.lib, args.func, args.outer)
    if len(sass) != len(body):
        print("warning: %d profiled instructions vs %d disassembled (library differs from the profiled one?)" % (len(body), len(sass)))
    n = min(len(sass), len(body))
    f = lambda r, k: float(r[col[k]] or 0)
    per_line = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0.0, 0])
    tot = [0.0, 0.0, 0.0, 0.0]
    for i in range(n):
        r = body[i]
        vals = (f(r, "Instructions Executed"), f(r, "Thread Instructions Executed"), f(r, "# Samples"), f(r, "stall_long_sb"))
        key = sass[i][2]
        for k in range(4):
            per_line[key][k] += vals[k]
            tot[k] += vals[k]
        per_line[key][4] += 1
    print("kernel total: %.3g warp instr, %.3g thread instr (%.1f threads/instr), %d samples, %d SASS instructions" % (
        tot[0], tot[1], tot[1] / max(tot[0], 1), tot[2], n))
    print("%-28s %6s %7s %7s %7s %8s" % ("file:line", "#sass", "%instr", "thr/in", "%sampl", "%long_sb"))
    for key, v in sorted(per_line.items(), key=lambda kv: -kv[1][0])[: args.top]:
        print("%-28s %6d %7.2f %7.1f %7.2f %8.2f" % ("%s:%d" % key, v[4], 100 * v[0] / tot[0], v[1] / max(v[0], 1),
                                                     100 * v[2] / max(tot[2], 1), 100 * v[3] / max(tot[2], 1)))
    if args.regions:
        regions = []
        for part in args.regions.split(","):
            nm, rng = part.split("=")
            lo, hi = rng.split("-")
            regions.append((nm, int(lo), int(hi)))
        agg = collections.OrderedDict((nm, [0.0, 0.0, 0.0, 0.0, 0]) for nm, _, _ in regions)
        agg["(other)"] = [0.0, 0.0, 0.0, 0.0, 0]
        for key, v in per_line.items():
            nm = next((nm for nm, lo, hi in regions if key[0].endswith(".cu") and lo <= key[1] <= hi), "(other)")
            for k in range(5):
                agg[nm][k] += v[k]
        print("\n%-28s %6s %7s %7s %7s %8s %9s" % ("region", "#sass", "%instr", "thr/in", "%sampl", "%long_sb", "%laneslot"))
        for nm, v in agg.items():
            print("%-28s %6d %7.2f %7.1f %7.2f %8.2f %9.2f" % (nm, v[4], 100 * v[0] / tot[0], v[1] / max(v[0], 1), 100 * v[2] / max(tot[2], 1),
                                                            100 * v[3] / max(tot[2], 1), 100 * v[1] / tot[1]))
    if args.sass:
        print("\nhottest SASS by samples")
        order = sorted(range(n), key=lambda i: -f(body[i], "# Samples"))[: args.top]
        for i in order:
            r = body[i]
            print("%5d %-26s %7.2f%% thr %5.1f  %s" % (i, "%s:%d" % sass[i][2], 100 * f(r, "# Samples") / max(tot[2], 1),
                                                     f(r, "Avg. Threads Executed"), r[col["Source"]].strip()[:90]))


if __name__ == "__main__":
    main()
