#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total and share.
usage: summarize_launches.py launches.csv [skip_first_n_launches]"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    tot = 0.0
    n = 0
    for r in rows[1:]:
        n += 1
        if n <= skip:
            continue
        name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").strip()
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(r[iu], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += v
        tot += v
    print(f"| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {c} | {t:.1f} | {100 * t / tot:.1f}% |")
    print(f"| **all** | {n - skip} | {tot:.1f} | 100% |")


if __name__ == "__main__":
    main()
