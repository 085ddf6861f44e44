"""Gauges: velocity at a point and surface water level along a line, computed on the device every step
and written by the host driver as GaugesVel_*.csv / GaugesSWL_*.csv in the reference's format.  tests/golden/gauges_ref/ holds
the files the reference CPU build wrote for the same case (computedt/outputdt 0.002 s; one SWL gauge on every step).
Compared row by row: same number of rows, times within 2e-5 relative (the dt tolerance), velocity within 5e-4 of max|v|+1e-6 m/s,
surface position within 1e-5 m (the kernel sums behind it are smooth in the particle state)."""
import json
import os
import subprocess

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
CLI = os.path.join(H.ROOT, "dualsphysics_b200", "lib", "b200sph")
REF = os.path.join(H.GOLDEN, "gauges_ref")


def _csv(path):
    lines = open(path).read().strip().splitlines()
    return lines[0], np.array([[float(x) for x in ln.split(";")] for ln in lines[1:]])


def test_gauges_match_reference(tmp_path, built_lib):
    fx = json.load(open(os.path.join(REF, "fixture_gauges.json")))
    H.write_case(tmp_path, **fx["spec"])
    out = tmp_path / "out"
    r = subprocess.run([CLI, str(tmp_path / "Case"), str(out)] + fx["options"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "Finished execution (code=0)" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    rep = {}
    for name in ("GaugesVel_Vel1.csv", "GaugesSWL_Swl1.csv", "GaugesSWL_SwlX.csv"):
        ho, o = _csv(str(out / name))
        hg, g = _csv(os.path.join(REF, name))
        assert ho == hg, (name, ho, hg)
        assert o.shape == g.shape, (name, o.shape, g.shape)
        assert np.all(np.abs(o[:, 0] - g[:, 0]) <= 2e-5 * np.maximum(g[:, 0], 1e-3) + 1e-6), name      # times (printed with 6 decimals)
        assert np.array_equal(o[:, 4:], g[:, 4:]), name                                                  # the configured points
        if "Vel" in name:
            err = float(np.max(np.abs(o[:, 1:4] - g[:, 1:4])))
            assert err < 5e-4 * float(np.max(np.abs(g[:, 1:4]))) + 1e-6, (name, err)
        else:
            err = float(np.max(np.abs(o[:, 1:4] - g[:, 1:4])))
            assert err < 1e-5, (name, err)
        rep[name] = {"rows": int(o.shape[0]), "max_abs_err": err}
    print("\n[gauges] " + json.dumps(rep))


REF_FORCE = os.path.join(H.GOLDEN, "gauges_force_ref")


def test_force_gauges_match_reference(tmp_path, built_lib):
    """Force gauges (JGaugeForce) on the fixed tank and on the moving piston against the CSV files the reference wrote
    (tools/make_gauges_force_fixture.py): same rows and times, force within 2e-3 of max|F| (pressure sums inherit the Tait EOS
    amplification of one ulp of rho)."""
    fx = json.load(open(os.path.join(REF_FORCE, "fixture_gauges.json")))
    H.write_case(tmp_path, **fx["spec"])
    out = tmp_path / "out"
    r = subprocess.run([CLI, str(tmp_path / "Case"), str(out)] + fx["options"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "Finished execution (code=0)" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    rep = {}
    for name in ("GaugesForce_Tank.csv", "GaugesForce_Piston.csv"):
        ho, o = _csv(str(out / name))
        hg, g = _csv(os.path.join(REF_FORCE, name))
        assert ho == hg, (name, ho, hg)
        assert o.shape == g.shape, (name, o.shape, g.shape)
        assert np.all(np.abs(o[:, 0] - g[:, 0]) <= 2e-5 * np.maximum(g[:, 0], 1e-3) + 1e-6), name
        err = float(np.max(np.abs(o[:, 1:] - g[:, 1:]))) / float(np.max(np.abs(g[:, 1:])))
        assert err < 2e-3, (name, err)
        rep[name] = {"rows": int(o.shape[0]), "rel_err": err}
    # maxz gauge of the same run: highest fluid particle near a vertical line (1e-6 m: float z of a particle position)
    ho, o = _csv(str(out / "GaugesMaxZ_Col.csv"))
    hg, g = _csv(os.path.join(REF_FORCE, "GaugesMaxZ_Col.csv"))
    assert ho == hg and o.shape == g.shape
    assert np.array_equal(o[:, 2:], g[:, 2:]) and float(np.max(np.abs(o[:, 1] - g[:, 1]))) < 1e-6
    rep["GaugesMaxZ_Col.csv"] = {"rows": int(o.shape[0]), "max_abs_err": float(np.max(np.abs(o[:, 1] - g[:, 1])))}
    print("\n[force gauges] " + json.dumps(rep))
