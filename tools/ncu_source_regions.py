#!/usr/bin/env python
"""Per-instruction view of one kernel of an .ncu-rep captured with --import-source on: executed warp instructions, active lanes
and warp-stall samples summed over contiguous SASS regions that run the same number of times (= the loops of the kernel), plus
the instructions with most stall samples.  Reads the report offline (no GPU):

    python tools/ncu_source_regions.py gpurun_out/r01c_forces_pair2_v5_10M.ncu-rep > profiles/r01c_source_regions_v5.txt
"""
import csv
import io
import subprocess
import sys

REASONS = ["stall_long_sb", "stall_wait", "stall_not_selected", "stall_selected", "stall_math", "stall_dispatch", "stall_short_sb",
           "stall_branch_resolving", "stall_no_inst", "stall_lg", "stall_mio", "stall_tex", "stall_barrier"]


def main(rep, mincount=100.0):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    kernel = rows[0][1]
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    ex = [int(r[ix["Instructions Executed"]]) for r in data]
    thr = [int(r[ix["Thread Instructions Executed"]]) for r in data]
    smp = [int(r[ix["# Samples"]]) for r in data]
    warps = ex[0]                                   # the first instruction runs once per warp
    tot, stot = sum(ex), sum(smp)
    print(f"kernel: {kernel}\nwarps: {warps}   SASS instructions: {len(data)}   executed per warp: {tot / warps:.0f}   "
          f"lanes active: {sum(thr) / tot:.1f}   stall samples: {stot}")
    print("\nregions (contiguous instructions with the same execution count within 15%; small ones merged into 'other'):")
    print(f"{'SASS index':>12} {'instr':>6} {'runs/warp':>10} {'% of executed':>14} {'% of samples':>13} {'lanes':>6}   dominant stall reasons (% of all samples)")
    seg = []
    for k, n in enumerate(ex):
        c = n / warps
        if seg and abs(seg[-1]["c"] - c) <= 0.15 * max(c, seg[-1]["c"], 1.0):
            s = seg[-1]
            s["b"] = k
        else:
            s = {"a": k, "b": k, "c": c, "ex": 0, "thr": 0, "smp": 0, "why": {r: 0 for r in REASONS}}
            seg.append(s)
        s["ex"] += n; s["thr"] += thr[k]; s["smp"] += smp[k]
        for r in REASONS:
            s["why"][r] += int(data[k][ix[r]])
    other = {"ex": 0, "smp": 0}
    wtot = sum(sum(s["why"].values()) for s in seg) or 1
    for s in seg:
        if s["ex"] / warps < mincount:
            other["ex"] += s["ex"]; other["smp"] += s["smp"]
            continue
        why = sorted(s["why"].items(), key=lambda kv: -kv[1])[:3]
        print(f"{s['a']:>5}-{s['b']:<6} {s['b'] - s['a'] + 1:>6} {s['c']:>10.1f} {100 * s['ex'] / tot:>13.1f}% {100 * s['smp'] / stot:>12.1f}% {s['thr'] / max(s['ex'], 1):>6.1f}   "
              + ", ".join(f"{r[6:]} {100 * v / wtot:.1f}" for r, v in why if v))
    print(f"{'other':>12} {'':>6} {'':>10} {100 * other['ex'] / tot:>13.1f}% {100 * other['smp'] / stot:>12.1f}%")
    print("\ninstructions with most stall samples:")
    for k in sorted(sorted(range(len(data)), key=lambda k: -smp[k])[:16]):
        print(f"{k:>6} {100 * smp[k] / stot:>6.2f}%  runs/warp {ex[k] / warps:>7.1f}  lanes {thr[k] / max(ex[k], 1):>4.1f}   {data[k][ix['Source']].strip()[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 100.0)
