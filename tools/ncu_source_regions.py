#!/usr/bin/env python
"""Per-instruction execution counts from an .ncu-rep source page, printed as a compact listing:
addr  warp-instr(M)  avg-threads  samples  SASS.  usage: ncu_source_regions.py rep [min_minstr]"""
import csv, io, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia, isrc, ismp, iex, ith = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Avg. Threads Executed')
tot = sum(float(r[iex]) for r in rows[2:] if len(r) > iex and r[iex])
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
print(f"total warp instructions: {tot/1e6:.1f} M")
for n, r in enumerate(rows[2:]):
    if len(r) <= iex or not r[iex]:
        continue
    ex = float(r[iex])
    if ex / 1e6 >= thr:
        print(f"{n:5d} {ex/1e6:9.2f} {float(r[ith] or 0):5.1f} {r[ismp]:>6s}  {r[isrc][:90]}")
