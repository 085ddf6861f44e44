"""GPU parity at BASELINE sizes against the COMPILED REFERENCE (tests/golden_large, tools/make_golden_large.py): the bench
configuration (Symplectic, Wendland, artificial viscosity, DDT2, DBC, -cellmode:half) at 170k particles (BASELINE configs[0] size)
and at 1M particles, every interaction variant.  Integer tables: bit-exact over the whole array through a position-weighted 64-bit
checksum.  Floats: the reference's values of every particle whose id is a multiple of the fixture's stride (~5.9k particles),
relative to max|field| of the subsample, with the measured per-field tolerances of tests/tolerances.py."""
import json
import os

import numpy as np
import pytest

import helpers as H
import tolerances as T
import dualsphysics_b200 as dsb
from dualsphysics_b200 import b2sd

pytestmark = pytest.mark.gpu
BASE = os.path.join(H.ROOT, "tests", "golden_large")
FIXTURES = sorted(d for d in os.listdir(BASE) if os.path.exists(os.path.join(BASE, d, "fixture.json"))) if os.path.isdir(BASE) else []


def _load(fx, tag):
    return b2sd.load(os.path.join(BASE, fx, tag + ".b2sd"))


def _hash(g, name):
    lo, hi = g["hash_" + name]
    return int(lo) | (int(hi) << 32)


def _tables(ctx, g, label):
    info = ctx.info()
    assert (info.np, info.npb, info.npbok) == (int(g["np"]), int(g["npb"]), int(g["npbok"])), label
    assert (info.ncx, info.ncy, info.ncz, info.nct) == (int(g["ncx"]), int(g["ncy"]), int(g["ncz"]), int(g["nct"])), label
    npb = int(g["npb"])
    for name in ("idp", "dcell", "code", "begincell"):
        a = ctx.array(name)
        assert a.size == int(g["len_" + name]) and H.table_hash(a) == _hash(g, name), f"{label}: {name} differs from the reference"
    if int(g["divide_full"]):
        for name in ("cellpart", "sortpart"):
            assert H.table_hash(ctx.array(name)[:info.np]) == _hash(g, name), f"{label}: {name}"
    return npb


def _subsample(ctx, g):
    """Indices (in the current sorted order) of the reference's subsample particles, in the subsample's order."""
    idp = ctx.array("idp")
    inv = np.empty(int(idp.max()) + 1, dtype=np.int64)
    inv[idp] = np.arange(idp.size)
    return inv[np.asarray(g["sub_idp"], np.int64)]


def _forces(ctx, g, p, key, tag):
    ix = _subsample(ctx, g)
    ace, ar, press = ctx.array("ace")[ix], ctx.array("ar")[ix], ctx.array("press")[ix]
    gace = g["sub_ace"].reshape(-1, 3)
    floor = float(p.rhopzero) * p.cs0 / float(p.kernelh) * 1e-4
    fluid = (ctx.array("code")[ix] & 0x1800) == 0x1800
    e = {"press": H.rel_err(press, g["sub_press"]), "ace": H.rel_err(ace[fluid], gace[fluid]),
         "ar": H.rel_err(ar, g["sub_ar"], floor=floor)}
    info = ctx.info()
    e["acemax"] = abs(info.acemax - float(g["acemax"])) / max(float(g["acemax"]), 1e-30)
    e["viscdtmax"] = abs(info.viscdtmax - float(g["viscdtmax"])) / max(float(g["viscdtmax"]), 1e-6)
    dtk = "dt_p" if "dt_p" in g else "dt"
    e["dt"] = abs(info.dt_last - float(g[dtk])) / float(g[dtk])
    H.parity_report(key, {"dump": tag, **e})
    fx = key.split("/")[0]
    for f in ("press", "ace", "ar", "acemax", "viscdtmax", "dt"):
        assert e[f] <= T.tol(fx, f), (key, tag, f, e[f], T.tol(fx, f))
    return e


@pytest.mark.parametrize("variant", [0, 5, 6])
@pytest.mark.parametrize("fx", FIXTURES)
def test_large_reference_parity(fx, variant, tmp_path, built_lib):
    meta = json.load(open(os.path.join(BASE, fx, "fixture.json")))
    H.write_case(tmp_path, **meta["spec"])
    opts = " ".join(meta["ref_options"])
    inter = dsb.lib.INTER_SYMPREDICTOR
    key = f"{fx}/variant{variant}"
    sim = dsb.Simulation(str(tmp_path / "Case"), options=opts)
    sim.ctx.set_variant(variant)
    p = sim.params
    rep = {}
    # step 0: interaction on the initial state (separate context: the interaction entry point may not be repeated while stepping)
    sim2 = dsb.Simulation(str(tmp_path / "Case"), options=opts)
    sim2.ctx.set_variant(variant)
    sim2.ctx.interaction(inter)
    rep["s0000_pred"] = _forces(sim2.ctx, _load(fx, "s0000_pred"), p, key, "s0000_pred")
    sim2.close()
    # steps 0..8, then the interaction of step 9 on a copy of the run
    sim3 = dsb.Simulation(str(tmp_path / "Case"), options=opts)
    sim3.ctx.set_variant(variant)
    sim3.ctx.run_steps(9)
    sim3.ctx.interaction(inter)
    rep["s0009_pred"] = _forces(sim3.ctx, _load(fx, "s0009_pred"), p, key, "s0009_pred")
    sim3.close()
    # ten whole steps: tables bit-exact, state of the subsample
    ctx = sim.ctx
    ctx.run_steps(10)
    ge = _load(fx, "s0009_end")
    _tables(ctx, ge, key + " end of step 9")
    ix = _subsample(ctx, ge)
    epos = float(np.max(np.abs(ctx.array("pos")[ix] - ge["sub_pos"].reshape(-1, 3))))
    evr = H.rel_err(ctx.array("velrho")[ix], ge["sub_velrho"].reshape(-1, 4))
    H.parity_report(key, {"dump": "s0009_end", "pos_abs_m": epos, "velrho": evr})
    rep["s0009_end"] = {"pos": epos, "velrho": evr}
    assert epos <= T.tol(fx, "pos_abs_m") and evr <= T.tol(fx, "velrho"), (key, epos, evr)
    assert abs(ctx.info().timestep - float(ge["timestep"])) <= 1e-6 * float(ge["timestep"])
    print(f"\n[large parity {key}] " + json.dumps(rep))
    sim.close()
