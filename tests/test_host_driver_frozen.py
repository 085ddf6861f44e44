"""The C++ host driver (csrc/host/sim.cpp: case parsing, run loop, PART schedule, every output file, restart) on the CPU: it is linked
here with tests/frozen_device/frozen_device.cpp, a stand-in for the device half of the C-ABI that keeps the uploaded particles as they
are and only advances the clock by 1/1024 s per step.  What the driver writes is compared with the files the reference CPU build wrote
for the same cases (tests/golden/ref_part, ref_part_float, ref_part_motion): the same item trees, value and array names and types; with a
frozen state the numbers that do not depend on the physics must be equal too (ids, initial positions, reference particles, PART
schedule).  The numerical content of a real run is what the `-m gpu` tests of tests/test_part_files.py check."""
import json
import os
import subprocess

import numpy as np
import pytest

import helpers as H
from dualsphysics_b200 import bi4

CSRC = os.path.join(H.ROOT, "dualsphysics_b200", "csrc")
DT = 1.0 / 1024.0


@pytest.fixture(scope="module")
def frozen(tmp_path_factory):
    """Builds the host driver + frozen device once: a shared library (for ctypes) and the command line on top of it."""
    d = tmp_path_factory.mktemp("frozen")
    src = [os.path.join(CSRC, "config.cpp")] + [os.path.join(CSRC, "host", f) for f in ("sim.cpp", "bi4.cpp", "xml_lite.cpp", "motionref.cpp")]
    src.append(os.path.join(H.ROOT, "tests", "frozen_device", "frozen_device.cpp"))
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")   # csrc/config.cpp sees vector types through b2s_internal.h
    if not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found (needed to compile csrc/config.cpp for the host)")
    subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(H.ROOT, "include"), "-I", cuda_inc, "-x", "c++"] + src
                   + ["-o", str(d / "libb200sph.so")], check=True)
    subprocess.run(["g++", "-O1", "-std=c++17", os.path.join(CSRC, "host", "main.cpp"), "-o", str(d / "b200sph_frozen"), "-L", str(d), "-lb200sph",
                    "-Wl,-rpath,$ORIGIN"], check=True)
    return {"exe": str(d / "b200sph_frozen"), "lib": str(d / "libb200sph.so")}


@pytest.fixture(scope="module")
def cli(frozen):
    return frozen["exe"]


def _structure(it, path=""):
    out = [(path + it.name, tuple((n, t) for n, t, _ in it.values), tuple((n, t) for n, t, _ in it.arrays))]
    for c in it.items:
        out += _structure(c, path + it.name + "/")
    return out


def _list_structure(items):
    return [(it.name, it.hide, tuple((n, t) for n, t, _ in it.values), tuple((n, t) for n, t, _ in it.arrays)) for it in items]


def _run(cli, case, out, *opts):
    r = subprocess.run([cli, str(case), str(out)] + list(opts), capture_output=True, text=True, timeout=120)
    return r


def test_part_files_and_head(cli, tmp_path):
    H.write_case(tmp_path, dp=0.05, params={"TimeMax": 0.01, "TimeOut": 0.002})
    out = tmp_path / "out"
    r = _run(cli, tmp_path / "Case", out, "-symplectic", "-tmax:0.004", "-tout:0.002")
    assert r.returncode == 0 and "Finished execution (code=0)" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
    ref = os.path.join(H.GOLDEN, "ref_part")
    for name in ("Part_0002.bi4", "Part_Head.ibi4"):
        so, sr = _structure(bi4.load(str(out / "data" / name))), _structure(bi4.load(os.path.join(ref, name)))
        assert so == sr, f"{name}: structure differs\nours={so}\nref ={sr}"
    # Part_Head.ibi4: every value and array of the reference's file (case constants, viscosity, MK blocks ...) except who wrote it and when
    def same_values(a, b, path=""):
        for (n, _, va), (_, _, vb) in zip(a.values, b.values):
            if n in ("AppName", "Date", "RunCode"):
                continue
            ok = np.allclose(va, vb, rtol=1e-6, atol=0) if isinstance(va, (float, tuple, list, np.ndarray)) else va == vb
            assert ok, (path + a.name, n, va, vb)
        for (n, _, va), (_, _, vb) in zip(a.arrays, b.arrays):
            assert np.allclose(np.asarray(va, np.float64), np.asarray(vb, np.float64), rtol=1e-6, atol=0), (path + a.name, n)
        for ca, cb in zip(a.items, b.items):
            same_values(ca, cb, path + a.name + "/")
    same_values(bi4.load(str(out / "data" / "Part_Head.ibi4")), bi4.load(os.path.join(ref, "Part_Head.ibi4")))
    # PART schedule (JDsOutputTime::GetNextTime): the first step whose time reaches 0.002, 0.004
    for k, tmin in ((0, 0.0), (1, 0.002), (2, 0.004)):
        p = bi4.load_part(str(out / "data" / f"Part_{k:04d}.bi4"))
        t = p["part"]["TimeStep"]
        assert tmin <= t < tmin + DT and p["part"]["Step"] == round(t / DT), (k, t)
    assert not (out / "data" / "Part_0003.bi4").exists()
    # frozen state: the particles of the case file come back, ordered by id, positions to the bit
    case = bi4.load_part(str(tmp_path / "Case.bi4"))
    io = np.argsort(case["Idp"], kind="stable")
    assert np.array_equal(p["Idp"], case["Idp"][io]) and np.array_equal(p["pos"], case["pos"][io])
    g = bi4.load_part(os.path.join(ref, "Part_0002.bi4"))
    assert len(p["Idp"]) == len(g["Idp"]) and np.array_equal(np.sort(g["Idp"]), p["Idp"])


def test_list_files_of_a_case_with_piston_and_floating_body(cli, tmp_path):
    spec = json.load(open(os.path.join(H.GOLDEN, "ref_part_motion", "fixture_motionref.json")))["spec"]
    H.write_case(tmp_path, **spec)
    out = tmp_path / "out"
    r = _run(cli, tmp_path / "Case", out)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    data = out / "data"
    reff, refm = os.path.join(H.GOLDEN, "ref_part_float"), os.path.join(H.GOLDEN, "ref_part_motion")
    for name, refdir in (("PartInfo.ibi4", reff), ("PartFloatInfo.ibi4", reff), ("PartMotionRef.ibi4", refm)):
        so, sr = _list_structure(bi4.load_list(str(data / name))), _list_structure(bi4.load_list(os.path.join(refdir, name)))
        assert so == sr, f"{name}: structure differs\nours={so}\nref ={sr}"
    # nothing is excluded by a frozen state: PartOut holds the head only, the same head as the reference's file
    oo, rr = bi4.load_list(str(data / "PartOut_000.obi4")), bi4.load_list(os.path.join(reff, "PartOut_000.obi4"))
    assert len(oo) == 1 and _list_structure(oo) == _list_structure(rr[:1])
    # reference particles and their positions: the bodies do not move here, so every PART repeats the head positions
    om, rm = bi4.load_list(str(data / "PartMotionRef.ibi4")), bi4.load_list(os.path.join(refm, "PartMotionRef.ibi4"))
    for name in ("MkBound", "Nid", "Id", "Ps"):
        assert np.array_equal(om[0].array(name), rm[0].array(name)), name
    for name in ("TimeOut", "MkBoundFirst", "MkMovingCount", "MkFloatCount", "MkCount", "FormatVer", "MainFile"):
        assert om[0].get(name) == rm[0].get(name), name
    for it in om[1:]:
        assert np.array_equal(it.array("PosRef"), om[0].array("Ps"))
    # floating-body file: constant head data equal to the reference's, one item per PART
    of, rf = bi4.load_list(str(data / "PartFloatInfo.ibi4")), bi4.load_list(os.path.join(reff, "PartFloatInfo.ibi4"))
    for name in ("MkBound", "Beginp", "Countp", "Mass", "Massp"):
        assert np.array_equal(of[0].array(name), rf[0].array(name)), name
    assert [it.get("Cpart") for it in of[1:]] == [it.get("Cpart") for it in rf[1:]] == list(range(5))
    assert np.max(np.abs(of[1].array("center") - rf[1].array("center"))) < 1e-12      # PART 0: the initial centre
    # PartInfo: counts of every PART
    oi = bi4.load_list(str(data / "PartInfo.ibi4"))
    for k, it in enumerate(oi[1:]):
        assert it.get("Cpart") == k and it.get("Npok") == 1650 and it.get("Nout") == 0 and it.get("npbin") == 1182 + 192
        assert k == 0 or abs(it.get("dtmean") - DT) < 1e-15


def test_restart_continues_numbering_time_and_list_files(cli, tmp_path):
    spec = json.load(open(os.path.join(H.GOLDEN, "ref_part_motion", "fixture_motionref.json")))["spec"]
    H.write_case(tmp_path, **spec)
    a, b = tmp_path / "runA", tmp_path / "runB"
    assert _run(cli, tmp_path / "Case", a, "-symplectic").returncode == 0
    r = _run(cli, tmp_path / "Case", b, "-symplectic", "-partbegin:2", str(a / "data"))
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert sorted(f for f in os.listdir(b / "data") if f.startswith("Part_0")) == ["Part_0002.bi4", "Part_0003.bi4", "Part_0004.bi4"]
    pa, pb = bi4.load_part(str(a / "data" / "Part_0002.bi4")), bi4.load_part(str(b / "data" / "Part_0002.bi4"))
    assert pb["part"]["TimeStep"] == pa["part"]["TimeStep"] and np.array_equal(pa["pos"], pb["pos"]) and np.array_equal(pa["Idp"], pb["Idp"])
    assert bi4.load_part(str(b / "data" / "Part_0004.bi4"))["part"]["TimeStep"] == bi4.load_part(str(a / "data" / "Part_0004.bi4"))["part"]["TimeStep"]
    # the floating body restarts from PartFloatInfo.ibi4 of the first run; the list files of the restart hold PARTs 2..4
    fb = bi4.load_list(str(b / "data" / "PartFloatInfo.ibi4"))
    assert [it.get("Cpart") for it in fb[1:]] == [2, 3, 4]
    mb = bi4.load_list(str(b / "data" / "PartMotionRef.ibi4"))
    assert [it.get("Cpart") for it in mb[1:]] == [2, 3, 4]
    # errors of the restart are loud
    r = _run(cli, tmp_path / "Case", b, "-symplectic", "-partbegin:2")
    assert r.returncode != 0 and "partbegin" in (r.stdout + r.stderr)
    r = _run(cli, tmp_path / "Case", b, "-symplectic", "-partbegin:9", str(a / "data"))
    assert r.returncode != 0


def test_gauge_files(cli, tmp_path):
    fx = json.load(open(os.path.join(H.GOLDEN, "gauges_force_ref", "fixture_gauges.json")))
    H.write_case(tmp_path, **fx["spec"])
    out = tmp_path / "out"
    r = _run(cli, tmp_path / "Case", out, *fx["options"])
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    refdir = os.path.join(H.GOLDEN, "gauges_force_ref")
    for name in ("GaugesForce_Tank.csv", "GaugesForce_Piston.csv", "GaugesMaxZ_Col.csv"):
        ours, ref = open(out / name).read().splitlines(), open(os.path.join(refdir, name)).read().splitlines()
        assert ours[0] == ref[0], name                                     # the reference's header line
        assert len(ours) > 2 and all(len(ln.split(";")) == len(ref[1].split(";")) for ln in ours[1:]), name
        t = np.array([float(ln.split(";")[0]) for ln in ours[1:]])
        assert t[0] == 0.0 and np.all(np.diff(t) > 0)


def test_every_accepted_configuration_loads_and_steps(cli, tmp_path):
    """The configurations tests/test_host_config.py expects to be accepted, taken past the point that test can reach without a
    device: upload, registration of motions / damping zones / gauges / floating bodies, three steps, one PART."""
    from test_host_config import ACCEPTED
    for k, (spec, options) in enumerate(ACCEPTED):
        d = tmp_path / f"c{k}"
        d.mkdir()
        H.write_case(d, dp=0.05, **spec)
        r = _run(cli, d / "Case", d / "out", "-nsteps:3", *options.split())
        assert r.returncode == 0 and "Finished execution (code=0)" in r.stdout, (k, r.stdout[-800:] + r.stderr[-800:])
        assert (d / "out" / "data" / "Part_0000.bi4").exists() and "Steps of simulation..............: 3" in r.stdout
        # periodic axes reach the heads of the PART files (TpPeri = axis bit mask, increment = minus the domain size on its axis)
        mode = sum({"x": 1, "y": 2, "z": 4}[c] for c in spec.get("open_axes", ""))
        for name, root in (("Part_Head.ibi4", None), ("Part_0000.bi4", None)):
            root = bi4.load(str(d / "out" / "data" / name))
            assert root.get("PeriMode") == mode, (k, name)
            if mode & 2:
                inc = root.get("PeriYinc")
                assert inc[1] < 0 and abs(inc[1] + (root.get("MapPosMax")[1] - root.get("MapPosMin")[1])) < 1e-12


def test_slab_loading_partitions_the_case(frozen, tmp_path):
    """b2s_sim_load_slab (one rank of a multi-GPU run): every rank derives the same cut planes from the particle histogram, the
    slabs are disjoint, together they hold every particle of the case, and a rank holds exactly the particles of its cell planes."""
    import ctypes as C
    from dualsphysics_b200.lib import Params
    lib = C.CDLL(frozen["lib"])
    lib.b2s_host_last_error.restype = C.c_char_p
    H.write_case(tmp_path, dp=0.03, piston=True)
    case = str(tmp_path / "Case").encode()
    allp = bi4.load_part(str(tmp_path / "Case.bi4"))
    ntot = len(allp["Idp"])
    for nranks in (1, 2, 3, 4):
        seen, cuts0 = [], None
        for rank in range(nranks):
            sim = C.c_void_p()
            r = lib.b2s_sim_load_slab(case, b"-symplectic", 0, rank, nranks, C.byref(sim))
            assert r == 0, lib.b2s_host_last_error().decode()
            cuts, mx = (C.c_uint32 * (nranks + 1))(), C.c_uint32()
            assert lib.b2s_sim_slab_info(sim, cuts, nranks + 1, C.byref(mx)) == 0
            cuts = list(cuts)
            prm = Params()
            assert lib.b2s_sim_params(sim, C.byref(prm)) == 0
            assert cuts[0] == 0 and cuts[-1] == prm.domcells[0] and all(b > a for a, b in zip(cuts, cuts[1:]))
            cuts0 = cuts0 or cuts
            assert cuts == cuts0                                     # the same planes on every rank
            ctx = C.c_void_p()
            assert lib.b2s_sim_ctx(sim, C.byref(ctx)) == 0
            n = C.c_uint32()
            idp, pos = np.zeros(ntot, np.uint32), np.zeros((ntot, 3))
            vel, rho = np.zeros((ntot, 3), np.float32), np.zeros(ntot, np.float32)
            assert lib.b2s_download_particles_byid(ctx, C.byref(n), idp.ctypes.data_as(C.c_void_p), pos.ctypes.data_as(C.c_void_p),
                                                   vel.ctypes.data_as(C.c_void_p), rho.ctypes.data_as(C.c_void_p)) == 0
            n = n.value
            cx = np.clip(np.floor((pos[:n, 0] - prm.domposmin[0]) / float(prm.scell)).astype(np.int64), 0, prm.domcells[0] - 1)
            assert n > 0 and np.all((cx >= cuts[rank]) & (cx < cuts[rank + 1])), (nranks, rank)
            assert mx.value >= np.bincount(cx).max()
            seen.append(idp[:n].copy())
            lib.b2s_sim_free(sim)
        ids = np.concatenate(seen)
        assert len(ids) == ntot and np.array_equal(np.sort(ids), np.sort(allp["Idp"])), nranks
        if nranks > 1:       # balanced within the granularity of one plane of cells
            counts = np.array([len(x) for x in seen])
            assert counts.max() <= ntot / nranks + mx.value


@pytest.mark.parametrize("refdir", ["gauges_ref", "gauges_force_ref"])
def test_gauge_schedule_matches_the_restated_rules(cli, tmp_path, refdir):
    """computedt / outputdt / start / end of every gauge (JGaugeItem::Update / Output): the rows the driver writes for the frozen clock
    t = k/1024 against the oracle's restatement of the same rules (oracle/oracle.py GaugeOracle, itself pinned to the reference's
    CSV files by tests/test_oracle_gauges.py) fed with the same times."""
    import oracle.oracle as orc
    fx = json.load(open(os.path.join(H.GOLDEN, refdir, "fixture_gauges.json")))
    H.write_case(tmp_path, **fx["spec"])
    out = tmp_path / "out"
    r = _run(cli, tmp_path / "Case", out, *fx["options"])
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    p, a = H.params_from_case(str(tmp_path / "Case"), {"StepAlgorithm": 2} if "-symplectic" in fx["options"] else None)
    go = orc.GaugeOracle(a["gauges_xml"], p, a["blocks"], a["timemax"], a["timeout"])
    go.measure = lambda sim, g: [0.0]
    t, k = 0.0, 0
    go.update(None, 0.0)
    while t < a["timemax"]:
        k += 1
        t = k * DT
        go.update(None, t)
    prefix = {"velocity": "GaugesVel_", "swl": "GaugesSWL_", "maxz": "GaugesMaxZ_", "force": "GaugesForce_"}
    assert go.items
    for g in go.items:
        if not g["out"]:
            continue
        rows = open(out / (prefix[g["type"]] + g["name"] + ".csv")).read().splitlines()[1:]
        ours = [float(ln.split(";")[0]) for ln in rows]
        want = [row[0] for row in go.rows[g["name"]]]
        assert len(ours) == len(want) and np.allclose(ours, want, atol=6e-7, rtol=0), (g["name"], ours[:5], want[:5])


def test_mdbc_restart_goes_through_the_extra_data_file(cli, tmp_path):
    """-svextraparts writes PartExtra_XXXX.bi4 like the reference (tests/golden/ref_part_extra/: same item, values, array; for fixed
    boundaries the same normals, i.e. twice the case's), and a restart with mDBC takes its normals from that file -- undoubled, so a
    second generation of the file is identical to the first -- or stops with the reference's message when the file is missing."""
    ref = os.path.join(H.GOLDEN, "ref_part_extra")
    fx = json.load(open(os.path.join(ref, "fixture_extra.json")))
    H.write_case(tmp_path, **fx["spec"])
    a, b, c = tmp_path / "runA", tmp_path / "runB", tmp_path / "runC"
    r = _run(cli, tmp_path / "Case", a, *fx["options"])
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert not (a / "data" / "PartExtra_0000.bi4").exists()
    ours, theirs = bi4.load(str(a / "data" / "PartExtra_0001.bi4")), bi4.load(os.path.join(ref, "PartExtra_0001.bi4"))
    assert _structure(ours) == _structure(theirs)
    for name in ("FormatVer", "CaseNbound", "CaseNfloat", "Cpart", "UseNormalsFt"):
        assert ours.get(name) == theirs.get(name), name
    assert np.array_equal(ours.array("Normals"), theirs.array("Normals"))
    # restart from PART 1 with the extra data, again saving extra data: PART 2 of the restart holds the same normals
    r = _run(cli, tmp_path / "Case", b, "-partbegin:1", str(a / "data"), *fx["options"])
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    again = bi4.load(str(b / "data" / "PartExtra_0002.bi4"))
    assert np.array_equal(again.array("Normals"), ours.array("Normals")) and again.get("Cpart") == 2
    # without the file: the reference's error
    os.remove(a / "data" / "PartExtra_0001.bi4")
    r = _run(cli, tmp_path / "Case", c, "-partbegin:1", str(a / "data"))
    assert r.returncode != 0 and "No extra data available to restart at PART_0001 with mDBC." in (r.stdout + r.stderr)
    # the list form of the option
    r = _run(cli, tmp_path / "Case", c, "-svextraparts:2,5-7")
    assert r.returncode == 0 and sorted(f for f in os.listdir(c / "data") if f.startswith("PartExtra")) == ["PartExtra_0002.bi4"]


def test_slab_ranks_write_part_pieces(frozen, tmp_path):
    """PART output of a multi-GPU run (b2s_sim_set_dirout + b2s_sim_save_part on every rank): one piece per rank in the reference's piece
    format (JPartDataBi4.cpp:99-127: Part_pRR_XXXX.bi4 with Piece / Npiece in the head, PartInfo_pRR.ibi4), Part_Head.ibi4 from rank 0,
    and the pieces of a PART hold every particle of the case exactly once."""
    import ctypes as C
    lib = C.CDLL(frozen["lib"])
    lib.b2s_host_last_error.restype = C.c_char_p
    H.write_case(tmp_path, dp=0.04, floating_box=(0.15, 0.25, 0.2, 0.1, 0.15, 0.1), floating_rel_density=0.5)
    case, out = str(tmp_path / "Case").encode(), tmp_path / "out"
    allp = bi4.load_part(str(tmp_path / "Case.bi4"))
    nranks = 3
    sims = []
    for rank in range(nranks):
        sim = C.c_void_p()
        assert lib.b2s_sim_load_slab(case, b"-symplectic", 0, rank, nranks, C.byref(sim)) == 0, lib.b2s_host_last_error().decode()
        assert lib.b2s_sim_set_dirout(sim, str(out).encode()) == 0
        sims.append(sim)
    for cpart in range(2):
        for sim in sims:
            assert lib.b2s_sim_save_part(sim) == 0, lib.b2s_host_last_error().decode()
    files = sorted(os.listdir(out / "data"))
    assert files == sorted([f"Part_p{r:02d}_{c:04d}.bi4" for r in range(nranks) for c in range(2)] + [f"PartInfo_p{r:02d}.ibi4" for r in range(nranks)]
                           + [f"PartOut_p{r:02d}_000.obi4" for r in range(nranks)] + ["Part_Head.ibi4", "PartFloatInfo.ibi4"])
    # excluded particles: one list file per piece (JPartOutBi4Save::GetFileNamePart, JPartOutBi4Save.cpp:102-113) with the reference's head
    refout = bi4.load_list(os.path.join(H.GOLDEN, "ref_part_float", "PartOut_000.obi4"))[0]
    for r in range(nranks):
        oo = bi4.load_list(str(out / "data" / f"PartOut_p{r:02d}_000.obi4"))
        assert len(oo) == 1 and _list_structure(oo) == _list_structure([refout])       # a frozen state excludes nothing: the head only
        assert oo[0].get("Piece") == r and oo[0].get("Npiece") == nranks
    ref = bi4.load(os.path.join(H.GOLDEN, "ref_part", "Part_0002.bi4"))
    ids = []
    for r in range(nranks):
        root = bi4.load(str(out / "data" / f"Part_p{r:02d}_0001.bi4"))
        assert root.get("Piece") == r and root.get("Npiece") == nranks and root.get("CaseNp") == len(allp["Idp"])
        assert [(n, t) for n, t, _ in root.values] == [(n, t) for n, t, _ in ref.values]            # the reference's head
        part = root.item("PART_0001")
        assert [(n, t) for n, t, _ in part.arrays] == [(n, t) for n, t, _ in ref.item("PART_0002").arrays]
        assert part.get("Npok") == len(part.array("Idp")) > 0
        ids.append(np.asarray(part.array("Idp")))
    ids = np.concatenate(ids)
    assert len(ids) == len(allp["Idp"]) and np.array_equal(np.sort(ids), np.sort(allp["Idp"]))
    assert bi4.load(str(out / "data" / "Part_Head.ibi4")).get("Npiece") == nranks
    for sim in sims:
        lib.b2s_sim_free(sim)


def test_extra_frequency_body_files(cli, tmp_path):
    """-toutx / TimeOutExtra: PartFloatInfo2.ibi4 and PartMotionRef2.ibi4 (JSph.cpp:789-790,859-870, JSphCpuSingle.cpp:1201-1206,
    JPartMotRefBi4Save.cpp:225-278).  The records a PART item holds follow from the two schedules alone, so the frozen clock pins
    them: a step whose end reaches the extra schedule adds a record to the PART that is written next; one record is stored like a
    PART of the main file, several as arrays.  Heads and item structures equal the reference's files
    (tests/golden/ref_part_extra2, tools/make_ref_part_extra2.py)."""
    refdir = os.path.join(H.GOLDEN, "ref_part_extra2")
    fx = json.load(open(os.path.join(refdir, "fixture_extra2.json")))
    H.write_case(tmp_path, **fx["spec"])
    out = tmp_path / "out"
    r = _run(cli, tmp_path / "Case", out, "-toutx:%g" % fx["toutx"])
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    # the schedule restated: t_n = n/1024; extra records first, then the PART of the same step (JSphCpuSingle.cpp:1201-1230)
    tmax, tout, toutx = fx["spec"]["params"]["TimeMax"], fx["spec"]["params"]["TimeOut"], fx["toutx"]

    def nxt(t, period):
        ct = int(t / period)
        tn = period * ct
        while tn <= t:
            tn = period * ct
            ct += 1
        return tn
    expect, pending, part = {0: [(0.0, 0)]}, [], 1
    n, t, nx, npart = 0, 0.0, nxt(0.0, toutx), nxt(0.0, tout)
    while t < tmax:
        n += 1
        t = n * DT
        if t >= nx:
            pending.append((t, n)); nx = nxt(t, toutx)
        if t >= npart:
            if pending:
                expect[part] = pending
            pending, part, npart = [], part + 1, nxt(t, tout)
    for name, arr in (("PartFloatInfo2.ibi4", "center"), ("PartMotionRef2.ibi4", "PosRef")):
        ours, ref = bi4.load_list(str(out / "data" / name)), bi4.load_list(os.path.join(refdir, name))
        assert _list_structure(ours[:1]) == _list_structure(ref[:1]), name
        for key in ("FormatVer", "MainFile", "TimeOut", "MkBoundFirst"):
            assert ours[0].get(key) == ref[0].get(key), (name, key)
        assert ours[0].get("MainFile") is False or ours[0].get("MainFile") == 0
        single = next(_list_structure([it])[0][2:] for it in ref[1:] if it.get("Count") is None)
        multi = next(_list_structure([it])[0][2:] for it in ref[1:] if it.get("Count") is not None)
        assert [it.get("Cpart") for it in ours[1:]] == sorted(expect), name
        per = 1 if arr == "center" else 6                      # one floating body; two bodies x three reference particles
        for it in ours[1:]:
            rec = expect[it.get("Cpart")]
            if len(rec) == 1:
                assert _list_structure([it])[0][2:] == single and it.get("Step") == rec[0][1] and abs(it.get("TimeStep") - rec[0][0]) < 1e-15
            else:
                assert _list_structure([it])[0][2:] == multi and it.get("Count") == len(rec)
                assert list(it.array("Step")) == [s for _, s in rec] and np.allclose(it.array("TimeStep"), [x for x, _ in rec], atol=1e-15)
            assert np.asarray(it.array(arr)).reshape(-1, 3).shape[0] == per * len(rec)
    # frozen bodies: every record repeats the reference positions of the head
    om = bi4.load_list(str(out / "data" / "PartMotionRef2.ibi4"))
    for it in om[1:]:
        pr = np.asarray(it.array("PosRef")).reshape(-1, 6, 3)
        assert all(np.array_equal(q, np.asarray(om[0].array("Ps")).reshape(6, 3)) for q in pr)
    assert os.path.exists(out / "data" / "PartFloatInfo.ibi4") and os.path.exists(out / "data" / "PartMotionRef.ibi4")   # the main files stay
    # the same schedule from the case file (<parameter key="TimeOutExtra">, JSph.cpp:789) instead of the switch: identical files
    spec2 = dict(fx["spec"]); spec2["params"] = dict(spec2["params"], TimeOutExtra=fx["toutx"])
    (tmp_path / "b").mkdir()
    H.write_case(tmp_path / "b", **spec2)
    r2 = _run(cli, tmp_path / "b" / "Case", tmp_path / "out_b")
    assert r2.returncode == 0, r2.stdout[-1500:] + r2.stderr[-1500:]
    for name in ("PartFloatInfo2.ibi4", "PartMotionRef2.ibi4"):
        assert open(out / "data" / name, "rb").read() == open(tmp_path / "out_b" / "data" / name, "rb").read(), name
    # without the option nothing extra is written
    r3 = _run(cli, tmp_path / "Case", tmp_path / "out_c")
    assert r3.returncode == 0 and not os.path.exists(tmp_path / "out_c" / "data" / "PartFloatInfo2.ibi4")
    # slab-decomposed runs refuse the option instead of dropping the files
    import ctypes as C
    lib = C.CDLL(os.path.join(os.path.dirname(cli), "libb200sph.so"))
    lib.b2s_host_last_error.restype = C.c_char_p
    sim = C.c_void_p()
    assert lib.b2s_sim_load_slab(str(tmp_path / "Case").encode(), b"-toutx:0.002", 0, 0, 2, C.byref(sim)) != 0
    assert b"toutx" in lib.b2s_host_last_error()


def _spawn_ranks(cli, case, out, world, ranks, idfile, opts=(), env_extra=None, rank_opts=None):
    procs = []
    for r in ranks:
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), B2S_NCCL_ID_FILE=idfile, **(env_extra or {}))
        o = list(opts) + list((rank_opts or {}).get(r, ()))
        procs.append(subprocess.Popen([cli, str(case), str(out)] + o, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    return procs


def test_multi_rank_start_roll_call(cli, tmp_path):
    """Start of a multi-process run (csrc/host/main.cpp): after loading its slab every rank answers a roll call in files next to the
    NCCL id, so that a rank that could not start (no such device, bad switch) stops its peers instead of leaving them inside the
    communicator's initialisation for ever, and a rank that never shows up is given up after B2S_START_TIMEOUT seconds."""
    import time
    H.write_case(tmp_path, dp=0.05, params={"TimeMax": 0.01, "TimeOut": 0.005})
    case = tmp_path / "Case"
    # 1. both ranks come up: the run goes through and writes its pieces
    idf = str(tmp_path / "id1")
    procs = _spawn_ranks(cli, case, tmp_path / "o1", 2, (0, 1), idf)
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
    assert "Finished execution (code=0)" in outs[0]
    assert os.path.exists(tmp_path / "o1" / "data" / "Part_p01_0001.bi4")
    assert not any(os.path.exists(idf + s) for s in ("", ".r0", ".r1"))           # nothing is left behind
    # 2. rank 1 cannot load (an unsupported switch): rank 0 stops at once with its reason
    idf = str(tmp_path / "id2")
    t0 = time.time()
    procs = _spawn_ranks(cli, case, tmp_path / "o2", 2, (0, 1), idf, rank_opts={1: ["-shifting:fulladv"]})
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert procs[0].returncode == 1 and procs[1].returncode == 1, outs
    assert "rank 1 could not start" in outs[0] and time.time() - t0 < 60
    # 3. rank 1 never shows up: rank 0 gives up after the limit
    idf = str(tmp_path / "id3")
    t0 = time.time()
    procs = _spawn_ranks(cli, case, tmp_path / "o3", 2, (0,), idf, env_extra={"B2S_START_TIMEOUT": "2"})
    out0 = procs[0].communicate(timeout=120)[0]
    assert procs[0].returncode == 1 and "not every rank of 2 came up" in out0 and time.time() - t0 < 60, out0
    # a later run with the same id file is not confused by what case 3 left behind
    procs = _spawn_ranks(cli, case, tmp_path / "o4", 2, (0, 1), idf)
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)
