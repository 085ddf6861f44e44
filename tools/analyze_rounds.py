#!/usr/bin/env python
"""CPU-only analysis of the lane imbalance of interaction variant 4 (forces_pair2.cu) -- what the "rounds" of its packed pair body
would cost under different ways of grouping the neighbour rows, on REAL particle configurations from the C oracle.

Variant 4 walks, per target particle (lane), the 5x5 rows (y,z) of half cells; per row a lane collects its hits (r<2h) in a bit
mask and the warp then runs max_lanes ceil(hits/2) rounds of the two-hits-per-lane body.  ncu says 21-26 of 32 lanes are active
in those rounds.  This script counts hits per (lane, row) for warps of 32 consecutive fluid particles in the solver's sorted order
(cell-sorted, x fastest, cells of size h) and reports rounds per particle for:

  current      sum over rows of  max_lanes ceil(h_row/2)
  pair-opp     rows paired with their point mirror (y,z)<->(-y,-z), hits paired inside a row:  max_lanes (ceil(hA/2)+ceil(hB/2))
  pair-opp-mix same, hits paired across the two rows:  max_lanes ceil((hA+hB)/2)
  z-slab       all 5 rows of a z plane in one batch:   max_lanes ceil(sum h/2)
  all rows     every row in one batch (full per-lane decoupling):  max_lanes ceil(sum h / 2)
  ideal        mean over lanes of (sum h)/2 (no imbalance, no odd halves)

usage: python tools/analyze_rounds.py [dp] [steps]   (default dp=0.02 -> ~29k fluid particles; steps = oracle steps before sampling)"""
import os
import sys
import tempfile

import numpy as np
from scipy.spatial import cKDTree

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402


def analyse(pos_f, h, label):
    """pos_f: fluid positions (n,3); cells of size h (cellmode half); warps = 32 consecutive particles in (cz,cy,cx) order."""
    cell = np.floor((pos_f - pos_f.min(axis=0) + 1e-9) / h).astype(np.int64)
    order = np.lexsort((cell[:, 0], cell[:, 1], cell[:, 2]))     # x fastest, stable
    pos, cell = pos_f[order], cell[order]
    n = (len(pos) // 32) * 32
    tree = cKDTree(pos)
    pairs = tree.query_pairs(2.0 * h, output_type="ndarray")
    i = np.concatenate([pairs[:, 0], pairs[:, 1]]); j = np.concatenate([pairs[:, 1], pairs[:, 0]])
    dy, dz = cell[j, 1] - cell[i, 1], cell[j, 2] - cell[i, 2]
    ok = (np.abs(dy) <= 2) & (np.abs(dz) <= 2) & (i < n)
    row = (dz[ok] + 2) * 5 + (dy[ok] + 2)
    hits = np.zeros((n, 25), np.int64)
    np.add.at(hits, (i[ok], row), 1)
    w = hits.reshape(-1, 32, 25)                                   # warp, lane, row
    c2 = lambda a: (a + 1) // 2                                    # noqa: E731
    tot = w.sum(axis=2)
    res = {}
    res["current"] = c2(w).max(axis=1).sum(axis=1)
    mir = w[:, :, ::-1]                                            # row r <-> 24-r
    a, b = w[:, :, :12], mir[:, :, :12]
    res["pair-opp"] = (c2(a) + c2(b)).max(axis=1).sum(axis=1) + c2(w[:, :, 12]).max(axis=1)
    res["pair-opp-mix"] = c2(a + b).max(axis=1).sum(axis=1) + c2(w[:, :, 12]).max(axis=1)
    res["z-slab"] = c2(w.reshape(-1, 32, 5, 5).sum(axis=3)).max(axis=1).sum(axis=1)
    res["all rows"] = c2(tot).max(axis=1)
    ideal = tot.mean(axis=1) / 2.0
    print(f"{label}: {n} fluid particles, {len(w)} warps, hits per particle mean {tot.mean():.1f}")
    print(f"  {'grouping':14s} rounds/particle   vs current   lanes active in the body")
    for k, v in res.items():
        print(f"  {k:14s} {v.mean():8.1f}          {v.mean() / res['current'].mean():5.2f}        {32 * ideal.mean() / v.mean():5.1f} of 32")
    print(f"  {'ideal':14s} {ideal.mean():8.1f}          {ideal.mean() / res['current'].mean():5.2f}         32.0 of 32")


def main():
    dp = float(sys.argv[1]) if len(sys.argv) > 1 else 0.02
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    with tempfile.TemporaryDirectory() as td:
        H.write_case(td, dp=dp)
        ora = H.oracle_sim_from_case(os.path.join(td, "Case"))
        h = float(ora.c.kernelh)
        fl = (ora.code & 0x1800) == 0x1800
        analyse(ora.pos[fl].copy(), h, "lattice start (the state bench.py times after 3 warm-up steps)")
        for _ in range(steps):
            ora.step()
        fl = ((ora.code & 0x1800) == 0x1800) & ((ora.code & 0xe000) == 0)
        analyse(ora.pos[:ora.np_][fl[:ora.np_]].copy(), h, f"after {steps} steps (t={ora.timestep:.3f} s, disordered)")


if __name__ == "__main__":
    main()
