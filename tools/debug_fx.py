#!/usr/bin/env python
"""Debug helper: steps a golden fixture on the GPU and reports, per dumped step, how the state differs from the reference dump
(compared by particle id, so that an order difference does not hide the size of the state difference)."""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
import dualsphysics_b200 as dsb  # noqa: E402
from dualsphysics_b200 import b2sd  # noqa: E402

fx = sys.argv[1]
meta = json.load(open(os.path.join(H.GOLDEN, fx, "fixture.json")))
with tempfile.TemporaryDirectory() as td:
    H.write_case(td, **meta["spec"])
    sim = dsb.Simulation(os.path.join(td, "Case"), options=" ".join(meta["ref_options"]))
    ctx = sim.ctx
    for s in range(meta["nsteps"]):
        dt = ctx.compute_step()
        f = os.path.join(H.GOLDEN, fx, f"s{s:04d}_end.b2sd")
        if not os.path.exists(f):
            continue
        g = b2sd.load(f)
        idp, code, pos, vr, dcell = ctx.array("idp"), ctx.array("code"), ctx.array("pos"), ctx.array("velrho"), ctx.array("dcell")
        gpos, gvr = g["pos"].reshape(-1, 3), g["velrho"].reshape(-1, 4)
        print(f"step {s}: dt {dt:.9e} ref {float(g['stepdt']):.9e} np {len(idp)} ref {int(g['np'])} order_equal {np.array_equal(idp, g['idp'])}")
        own, gown = (code & 0xe000) == 0, (g["code"] & 0xe000) == 0
        o, go = np.argsort(idp[own], kind="stable"), np.argsort(g["idp"][gown], kind="stable")
        if np.array_equal(idp[own][o], g["idp"][gown][go]):
            dp_ = np.abs(pos[own][o] - gpos[gown][go]); dv = np.abs(vr[own][o] - gvr[gown][go])
            k = int(np.argmax(dp_.max(axis=1)))
            print(f"   by id: max|dpos| {dp_.max():.3e} at id {idp[own][o][k]} pos {pos[own][o][k]} ref {gpos[gown][go][k]}  max|dvel| {dv[:, :3].max():.3e} max|drho| {dv[:, 3].max():.3e}")
            dc = dcell[own][o] != g["dcell"][gown][go]
            print(f"   cells differing: {int(dc.sum())} ids {idp[own][o][dc][:10]}")
        else:
            print("   owned id sets differ", len(o), len(go))
        if len(idp) == len(g["idp"]):
            bad = np.nonzero(idp != g["idp"])[0]
            print(f"   order mismatches: {len(bad)} first at {bad[:6]} ours {idp[bad[:6]]} ref {g['idp'][bad[:6]]} codes {code[bad[:6]]} {g['code'][bad[:6]]}")
    sim.close()
