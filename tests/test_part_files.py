"""PART output interchange: the host driver must write Part_XXXX.bi4 / Part_Head.ibi4 that are
structurally identical to the reference's files (same items, value names and types, array names and types, in the same
order) and numerically equal within the step tolerances.  tests/golden/ref_part/ holds the reference CPU build's own
output for `write_case(dp=0.05, TimeMax 0.01, TimeOut 0.002)` run with `-symplectic -tmax:0.004 -tout:0.002`
(generated with oracle/_ref/DualSPHysics5.4CPU_linux64)."""
import os
import subprocess

import numpy as np
import pytest

import helpers as H
from dualsphysics_b200 import bi4

REF = os.path.join(H.GOLDEN, "ref_part")
CLI = os.path.join(H.ROOT, "dualsphysics_b200", "lib", "b200sph")


def _structure(it, path=""):
    out = [(path + it.name, "item", tuple((n, t) for n, t, _ in it.values), tuple((n, t) for n, t, _ in it.arrays))]
    for c in it.items:
        out += _structure(c, path + it.name + "/")
    return out


def test_reader_understands_reference_part_files():
    """Our JBinaryData reader against files written by the reference itself (no GPU needed)."""
    root = bi4.load(os.path.join(REF, "Part_0002.bi4"))
    assert root.name == "JPartDataBi4"
    part = root.item("PART_0002")
    assert part is not None and part.get("Npok") == len(part.array("Idp"))
    assert abs(part.get("TimeStep") - 0.004) < 2e-4
    d = bi4.load_part(os.path.join(REF, "Part_0002.bi4"))
    assert d["pos"].shape == (len(d["Idp"]), 3) and np.all(np.isfinite(d["pos"]))
    head = bi4.load(os.path.join(REF, "Part_Head.ibi4"))
    assert head.name == "JPartDataHead"


@pytest.mark.gpu
def test_driver_writes_reference_compatible_part_files(tmp_path, built_lib):
    H.write_case(tmp_path, dp=0.05, params={"TimeMax": 0.01, "TimeOut": 0.002})
    out = tmp_path / "out"
    r = subprocess.run([CLI, str(tmp_path / "Case"), str(out), "-symplectic", "-tmax:0.004", "-tout:0.002"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "Finished execution (code=0)" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    for name in ("Part_0000.bi4", "Part_0001.bi4", "Part_0002.bi4", "Part_Head.ibi4"):
        assert (out / "data" / name).exists(), name
    # ---- structure: identical item tree, value names/types and array names/types
    for name in ("Part_0002.bi4", "Part_Head.ibi4"):
        ours, ref = bi4.load(str(out / "data" / name)), bi4.load(os.path.join(REF, name))
        so, sr = _structure(ours), _structure(ref)
        assert so == sr, f"{name}: structure differs\nours={so}\nref ={sr}"
    # ---- data: same particles, same state within the step tolerances (pos 1e-7 m, vel/rho 2e-6 of max)
    o, g = bi4.load_part(str(out / "data" / "Part_0002.bi4")), bi4.load_part(os.path.join(REF, "Part_0002.bi4"))
    # (the reference stores particles in its cell-sorted order, this driver by id: compare by id)
    io, ig = np.argsort(o["Idp"], kind="stable"), np.argsort(g["Idp"], kind="stable")
    assert np.array_equal(o["Idp"][io], g["Idp"][ig])
    assert np.max(np.abs(o["pos"][io] - g["pos"][ig])) < 1e-7
    # velocity alone (max|v| = 0.04 m/s this early) inherits the ace tolerance (5e-4, see test_gpu_parity); rho 2e-6
    assert H.rel_err(o["Vel"][io], g["Vel"][ig]) < 5e-5 and H.rel_err(o["Rhop"][io], g["Rhop"][ig]) < 2e-6
    po, pg = bi4.load(str(out / "data" / "Part_0002.bi4")).item("PART_0002"), bi4.load(os.path.join(REF, "Part_0002.bi4")).item("PART_0002")
    assert abs(po.get("TimeStep") - pg.get("TimeStep")) <= 1e-6 * pg.get("TimeStep")
    assert po.get("Npok") == pg.get("Npok") and po.get("Step") == pg.get("Step")


@pytest.mark.gpu
def test_restart_from_part(tmp_path, built_lib):
    """-partbegin:N dir (JSphCfgRun.cpp:462-470, JSph.cpp:2214-2225): a run restarted from Part_0001.bi4 of a first run must
    arrive at the same state as the uninterrupted run.  The PART stores double positions and the float velocity/density the
    solver itself holds, plus SymplecticDtPre, so the restarted state is exact; only the particle order inside cells differs
    (the restart sorts by id), i.e. rounding-level differences: pos 1e-9 m, vel/rho 2e-6."""
    H.write_case(tmp_path, dp=0.04, params={"TimeMax": 0.03, "TimeOut": 0.01})
    case = str(tmp_path / "Case")
    a, b = tmp_path / "runA", tmp_path / "runB"
    r = subprocess.run([CLI, case, str(a), "-symplectic"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    r = subprocess.run([CLI, case, str(b), "-symplectic", "-partbegin:1", str(a / "data")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert not (b / "data" / "Part_0000.bi4").exists() and (b / "data" / "Part_0001.bi4").exists()
    for k in (1, 2, 3):
        pa, pb = bi4.load_part(str(a / "data" / f"Part_{k:04d}.bi4")), bi4.load_part(str(b / "data" / f"Part_{k:04d}.bi4"))
        assert np.array_equal(pa["Idp"], pb["Idp"])
        assert abs(pa["part"]["TimeStep"] - pb["part"]["TimeStep"]) <= 1e-9 * max(pa["part"]["TimeStep"], 1e-9), k
        assert np.max(np.abs(pa["pos"] - pb["pos"])) < 1e-9, k
        assert H.rel_err(pb["Vel"], pa["Vel"]) < 2e-6 and H.rel_err(pb["Rhop"], pa["Rhop"]) < 2e-6, k
    # a restart needs its directory argument, and PART files of another case are refused
    r = subprocess.run([CLI, case, str(b), "-symplectic", "-partbegin:1"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "partbegin" in (r.stdout + r.stderr)


REF_FLOAT = os.path.join(H.GOLDEN, "ref_part_float")
FLOAT_SPEC = dict(dp=0.05, floating_box=(0.15, 0.25, 0.2, 0.1, 0.15, 0.1), floating_rel_density=0.5, piston=True,
                  params={"TimeMax": 0.02, "TimeOut": 0.005, "RhopOutMin": 999.5})   # == tools/make_ref_part_float.py


def _list_structure(items):
    return [(it.name, it.hide, tuple((n, t) for n, t, _ in it.values), tuple((n, t) for n, t, _ in it.arrays)) for it in items]


def test_reader_understands_reference_list_files():
    """load_list against the reference's own PartInfo / PartOut / PartFloatInfo files (no GPU needed)."""
    info = bi4.load_list(os.path.join(REF_FLOAT, "PartInfo.ibi4"))
    assert info[0].name == "JPartDataBi4" and [it.name for it in info[1:]] == [f"PART_{k:04d}" for k in range(5)]
    out = bi4.load_list(os.path.join(REF_FLOAT, "PartOut_000.obi4"))
    assert out[0].name == "JPartOutBi4" and all(it.get("Nout") == len(it.array("Idp")) for it in out[1:])
    ft = bi4.load_list(os.path.join(REF_FLOAT, "PartFloatInfo.ibi4"))
    assert ft[0].get("FtCount") == 1 and len(ft) == 6 and ft[1].array("center").shape == (1, 3)


@pytest.mark.gpu
def test_driver_writes_reference_compatible_list_files(tmp_path, built_lib):
    """PartInfo.ibi4, PartOut_000.obi4 and PartFloatInfo.ibi4 against the files the reference wrote for the same
    case: same heads and items (names, value/array names and types), same excluded particles, same floating-body history."""
    H.write_case(tmp_path, **FLOAT_SPEC)
    out = tmp_path / "out"
    r = subprocess.run([CLI, str(tmp_path / "Case"), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "Finished execution (code=0)" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    for name in ("PartInfo.ibi4", "PartOut_000.obi4", "PartFloatInfo.ibi4"):
        ours, ref = bi4.load_list(str(out / "data" / name)), bi4.load_list(os.path.join(REF_FLOAT, name))
        so, sr = _list_structure(ours), _list_structure(ref)
        assert so == sr, f"{name}: structure differs\nours={so}\nref ={sr}"
    # excluded particles: the same ids per PART (density threshold; at most a few borderline particles may change PART)
    oo, rr = bi4.load_list(str(out / "data" / "PartOut_000.obi4")), bi4.load_list(os.path.join(REF_FLOAT, "PartOut_000.obi4"))
    alo, alr = set(), set()
    for a, b in zip(oo[1:], rr[1:]):
        assert a.get("Cpart") == b.get("Cpart") and abs(a.get("TimeStep") - b.get("TimeStep")) < 1e-6
        alo |= set(a.array("Idp").tolist()); alr |= set(b.array("Idp").tolist())
        assert set(np.unique(a.array("Motive")).tolist()) <= {1, 2, 3}
    assert len(alo ^ alr) <= 3, sorted(alo ^ alr)
    first_o, first_r = dict(zip(oo[1].array("Idp").tolist(), oo[1].array("Rhop"))), dict(zip(rr[1].array("Idp").tolist(), rr[1].array("Rhop")))
    common = sorted(set(first_o) & set(first_r))
    assert len(common) >= 0.9 * len(first_r) and max(abs(first_o[i] - first_r[i]) for i in common) < 2e-3   # rho ~1000: 2e-6 relative
    # PART info: counts of the interval
    io, ir = bi4.load_list(str(out / "data" / "PartInfo.ibi4")), bi4.load_list(os.path.join(REF_FLOAT, "PartInfo.ibi4"))
    for a, b in zip(io[1:], ir[1:]):
        assert a.get("Step") == b.get("Step") and abs(a.get("Nout") - b.get("Nout")) <= 3 and abs(a.get("noutrho") - b.get("noutrho")) <= 3
        assert abs(a.get("Npok") - b.get("Npok")) <= 3
    # floating body history
    fo, fr = bi4.load_list(str(out / "data" / "PartFloatInfo.ibi4")), bi4.load_list(os.path.join(REF_FLOAT, "PartFloatInfo.ibi4"))
    for name in ("MkBound", "Beginp", "Countp", "Mass", "Massp"):
        assert np.array_equal(fo[0].array(name), fr[0].array(name)), name
    # (the reference writes Radius=0 in the head: it configures the file before it computes the radius; ours holds the value)
    assert float(fo[0].array("Radius")[0]) > 0
    for a, b in zip(fo[1:], fr[1:]):
        assert a.get("Step") == b.get("Step")
        assert np.max(np.abs(a.array("center") - b.array("center"))) < 1e-7
        for name, tol in (("fvel", 1e-3), ("fomega", 1e-3), ("facelin", 5e-3), ("fluforcelin", 5e-3), ("fluforceang", 5e-3), ("preacelin", 5e-3)):
            ref = b.array(name).astype(np.float64)
            assert np.max(np.abs(a.array(name) - ref)) <= tol * max(np.max(np.abs(ref)), 1e-6), (name, a.get("Cpart"), a.array(name), ref)


@pytest.mark.gpu
def test_restart_with_floating_body(tmp_path, built_lib):
    """JSph::InitFloatingsRestart (JSph.cpp:2373-2391): the restart takes centre, velocities and radius of the body from
    PartFloatInfo.ibi4.  Like the reference it restarts with zero accumulated angles, so the inertia tensor is the initial one:
    the body has turned by ~1e-3 rad at the restart, which bounds the difference to the uninterrupted run."""
    spec = dict(FLOAT_SPEC); spec["params"] = {"TimeMax": 0.02, "TimeOut": 0.005}
    H.write_case(tmp_path, **spec)
    case = str(tmp_path / "Case")
    a, b = tmp_path / "runA", tmp_path / "runB"
    r = subprocess.run([CLI, case, str(a)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    r = subprocess.run([CLI, case, str(b), "-partbegin:2", str(a / "data")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    for k in (3, 4):
        pa, pb = bi4.load_part(str(a / "data" / f"Part_{k:04d}.bi4")), bi4.load_part(str(b / "data" / f"Part_{k:04d}.bi4"))
        assert np.array_equal(pa["Idp"], pb["Idp"])
        assert np.max(np.abs(pa["pos"] - pb["pos"])) < 1e-6, k
    fa, fb = bi4.load_list(str(a / "data" / "PartFloatInfo.ibi4")), bi4.load_list(str(b / "data" / "PartFloatInfo.ibi4"))
    ca = {it.get("Cpart"): it.array("center") for it in fa[1:]}
    cb = {it.get("Cpart"): it.array("center") for it in fb[1:]}
    assert set(cb) == {2, 3, 4} and all(np.max(np.abs(ca[k] - cb[k])) < 1e-6 for k in cb)


@pytest.mark.gpu
def test_driver_writes_motionref_file(tmp_path, built_lib):
    """PartMotionRef.ibi4 from a run of the driver against the reference's file for the same case
    (tests/golden/ref_part_motion/): same structure, the same three reference particles per body, their positions at every PART
    within the tolerance of the body trajectories (piston 1e-9 m, floating box 1e-7 m).  The writer itself is pinned bit for
    bit on the CPU by tests/test_motionref.py."""
    import json
    refdir = os.path.join(H.GOLDEN, "ref_part_motion")
    H.write_case(tmp_path, **json.load(open(os.path.join(refdir, "fixture_motionref.json")))["spec"])
    out = tmp_path / "out"
    r = subprocess.run([CLI, str(tmp_path / "Case"), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "Finished execution (code=0)" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    ours, ref = bi4.load_list(str(out / "data" / "PartMotionRef.ibi4")), bi4.load_list(os.path.join(refdir, "PartMotionRef.ibi4"))
    assert _list_structure(ours) == _list_structure(ref)
    for name in ("MkBound", "Nid", "Id", "Ps"):
        assert np.array_equal(ours[0].array(name), ref[0].array(name)), name
    for a, b in zip(ours[1:], ref[1:]):
        assert a.get("Cpart") == b.get("Cpart") and a.get("Step") == b.get("Step")
        assert np.max(np.abs(np.asarray(a.array("PosRef")) - np.asarray(b.array("PosRef")))) < 1e-7
    # Run.out (the log) and Run.csv (one-line summary with the reference's columns, JSph::GetResInfo)
    log = open(out / "Run.out").read()
    assert "Finished execution (code=0)" in log and "Steps of simulation" in log and "PART files.......................: 5" in log
    head, row = open(out / "Run.csv").read().strip().splitlines()
    assert head.startswith("#RunName;") and len(head.split(";")) == len(row.split(";")) == 26


@pytest.mark.gpu
def test_driver_writes_extra_frequency_files(tmp_path, built_lib):
    """-toutx (TimeOutExtra): PartFloatInfo2.ibi4 / PartMotionRef2.ibi4 from a run of the driver against the reference's files for
    the same case and switch (tests/golden/ref_part_extra2): same items (one per PART, a single record stored like a PART of the main
    file, several as arrays), the same steps in every record, reference-particle positions and floating-body centre within 1e-7 m,
    body velocities within 1e-3 of their maximum.  The positions come through b2s_get_pos_byid (device-side id table of the last
    divide); schedule and item layout are pinned on the CPU by tests/test_host_driver_frozen.py."""
    import json
    refdir = os.path.join(H.GOLDEN, "ref_part_extra2")
    fx = json.load(open(os.path.join(refdir, "fixture_extra2.json")))
    H.write_case(tmp_path, **fx["spec"])
    out = tmp_path / "out"
    r = subprocess.run([CLI, str(tmp_path / "Case"), str(out), "-toutx:%g" % fx["toutx"]], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "Finished execution (code=0)" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    for name, arrs in (("PartFloatInfo2.ibi4", (("center", 1e-7, False), ("fvel", 1e-3, True), ("fomega", 1e-3, True))),
                       ("PartMotionRef2.ibi4", (("PosRef", 1e-7, False),))):
        ours, ref = bi4.load_list(str(out / "data" / name)), bi4.load_list(os.path.join(refdir, name))
        assert _list_structure(ours) == _list_structure(ref), name
        for a, b in zip(ours[1:], ref[1:]):
            assert a.get("Cpart") == b.get("Cpart") and a.get("Count") == b.get("Count")
            if b.get("Count") is None:
                assert a.get("Step") == b.get("Step") and abs(a.get("TimeStep") - b.get("TimeStep")) <= 2e-6 * max(b.get("TimeStep"), 1e-3)
            else:
                assert np.array_equal(a.array("Step"), b.array("Step"))
                assert np.allclose(a.array("TimeStep"), b.array("TimeStep"), rtol=2e-6)
            for arr, tol, rel in arrs:
                x, y = np.asarray(a.array(arr), np.float64), np.asarray(b.array(arr), np.float64)
                scale = max(float(np.abs(y).max()), 1e-3) if rel else 1.0
                assert np.max(np.abs(x - y)) <= tol * scale, (name, a.name, arr, float(np.max(np.abs(x - y))), scale)
