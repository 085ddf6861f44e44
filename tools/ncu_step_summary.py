#!/usr/bin/env python
"""One line per kernel of an .ncu-rep of a whole step (tools/profile_step.py): launches, mean duration, DRAM bytes and achieved
DRAM GB/s, FP32 thread-level op counters and the FLOP/s they give, occupancy, issue-slot use.  usage: ncu_step_summary.py file.ncu-rep"""
import collections
import csv
import io
import re
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]


def col(name):
    return hdr.index(name) if name in hdr else None


def scale(i):          # to base units (bytes, ns -> s)
    u = units[i].lower()
    return {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}.get(u, 1.0)


C = {k: col(k) for k in ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                         "sm__sass_thread_inst_executed_op_fadd_pred_on.sum", "sm__sass_thread_inst_executed_op_fmul_pred_on.sum",
                         "sm__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
                         "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
                         "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"]}
agg = collections.OrderedDict()
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[C["Kernel Name"]]).replace("void ", "").strip()
    a = agg.setdefault(name, collections.defaultdict(float))
    a["n"] += 1

    def val(k):
        i = C[k]
        if i is None or not r[i]:
            return 0.0
        return float(r[i].replace(",", "")) * scale(i)
    a["t"] += val("gpu__time_duration.sum")
    a["rd"] += val("dram__bytes_read.sum"); a["wr"] += val("dram__bytes_write.sum")
    a["fadd"] += val("sm__sass_thread_inst_executed_op_fadd_pred_on.sum"); a["fmul"] += val("sm__sass_thread_inst_executed_op_fmul_pred_on.sum")
    a["ffma"] += val("sm__sass_thread_inst_executed_op_ffma_pred_on.sum")
    a["occ"] += val("sm__warps_active.avg.pct_of_peak_sustained_active"); a["issue"] += val("smsp__issue_active.avg.pct_of_peak_sustained_active")
    a["regs"] = val("launch__registers_per_thread")
    a["l1"] += val("l1tex__t_sector_hit_rate.pct"); a["l2"] += val("lts__t_sector_hit_rate.pct")
    a["l1sect"] += val("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")
tot = sum(a["t"] for a in agg.values())
print("| kernel | launches | us/launch | share | DRAM MB/launch (rd+wr) | DRAM GB/s | FP32 ops/launch (fadd+fmul+2*ffma, thread level) | GFLOP/s | occupancy % | issue slots % | regs | L1 hit % | L2 hit % | L1 global-load sectors/launch |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["t"]):
    n = a["n"]; t = a["t"] / n
    flop = (a["fadd"] + a["fmul"] + 2 * a["ffma"]) / n
    byt = (a["rd"] + a["wr"]) / n
    print(f"| `{name}` | {int(n)} | {t * 1e6:.1f} | {a['t'] / tot * 100:.1f}% | {a['rd'] / n / 1e6:.1f}+{a['wr'] / n / 1e6:.1f} | {byt / t / 1e9:.0f} | {flop:.3e} | {flop / t / 1e9:.0f} | "
          f"{a['occ'] / n:.0f} | {a['issue'] / n:.0f} | {int(a['regs'])} | {a['l1'] / n:.0f} | {a['l2'] / n:.0f} | {a['l1sect'] / n:.3e} |")
print(f"\nall kernels of the step: {tot * 1e3:.3f} ms under ncu (cold caches, serialised replays: shares, not absolutes)")
