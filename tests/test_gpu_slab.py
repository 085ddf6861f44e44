"""Slab decomposition on ONE GPU: N contexts in one process exchange halos/migrants through plain device
copies (dualsphysics_b200.slab.LocalTransport path) and must reproduce the single-context run.

The pair sums of a particle next to a cut visit its neighbours in a different order (ghosts are appended before the
stable sort), so floats agree to rounding, not bit for bit: pos 1e-7 m, vel/rho 2e-6 of max|field| after 10 steps (the
same tolerances the single-GPU path holds against the reference).  Particle SETS (ids) and the dt sequence are checked too.
"""
import numpy as np
import pytest

import helpers as H
import dualsphysics_b200 as dsb
from dualsphysics_b200 import slab

pytestmark = pytest.mark.gpu

CASES = [
    ("sym_art", 0.03, {}, "", 2),
    ("sym_art_3slabs", 0.03, {}, "", 3),
    ("sym_art_half", 0.03, {}, "-cellmode:half", 2),
    ("ver_cubic", 0.03, {"StepAlgorithm": 1, "Kernel": 1, "DensityDT": 1}, "", 2),
    ("sym_lamsps", 0.03, {"ViscoTreatment": 2, "Visco": 1e-6, "DensityDT": 3}, "", 2),
]


@pytest.mark.parametrize("native", [True, False], ids=["cxx_step", "py_step"])
@pytest.mark.parametrize("name,dp,opts,options,nranks", CASES, ids=[c[0] for c in CASES])
def test_slabs_match_single_context(name, dp, opts, options, nranks, native, tmp_path, built_lib):
    """native: the whole step inside the library (csrc/slab_comm.cu, device-counted messages, local test transport);
    otherwise the round-1 orchestration of slab.py (host-counted messages, LocalTransport)."""
    H.write_case(tmp_path, dp=dp, params=opts)
    case = str(tmp_path / "Case")
    ref = dsb.Simulation(case, options=options)
    ss = slab.SlabSimulation(case, options=options, nranks=nranks, dist=None, native=native)
    assert ss.native == (native and ss._merged_ok())      # slabs no wider than the halo keep the host-counted exchange
    assert len(ss.cuts) == nranks + 1 and all(b > a for a, b in zip(ss.cuts, ss.cuts[1:]))
    n0 = ref.ctx.info().np
    assert sum(ss.owned_counts()) == n0, "slabs do not partition the particle set"
    nsteps = 10
    for s in range(nsteps):
        dt_ref = ref.ctx.compute_step()
        dt = ss.step()
        assert abs(dt - dt_ref) <= 2e-5 * dt_ref, (s, dt, dt_ref)
    ids, pos, vr = ss.gather_byid()
    rid = ref.ctx.array("idp"); o = np.argsort(rid, kind="stable")
    rpos, rvr = ref.ctx.array("pos")[o], ref.ctx.array("velrho")[o]
    assert np.array_equal(ids, rid[o]), "particle set differs from the single-context run"
    epos = float(np.max(np.abs(pos - rpos)))
    evr = H.rel_err(vr, rvr)
    print(f"\n[slab {name} x{nranks}] cuts={ss.cuts} owned={ss.owned_counts()} pos_err={epos:.3e} velrho_err={evr:.3e} "
          f"exchanged={ss.exchanged_bytes / nsteps / 1e3:.1f} kB/step")
    assert epos < 1e-7 and evr < 2e-6
    # load balance of the cuts: a whole cell plane is the granularity, so the densest plane bounds the imbalance
    oc = np.array(ss.owned_counts(), np.float64)
    p = ref.params
    planes = np.bincount(ref.ctx.array("dcell") >> ((int(p.domcellcode) >> 10) & 31))
    assert oc.max() <= oc.mean() + planes.max()
    ss.close(); ref.close()


@pytest.mark.parametrize("options", ["-symplectic", "-verlet", "-symplectic -viscolamsps:1e-6 -ddt:3"])
def test_slab_migration(options, tmp_path, built_lib):
    """Long enough for the collapsing column to push particles across the cuts: migrants carry the full integrator state
    (predictor copies / Verlet m1 / SPS tau).  A migrant is appended at the end of its new cell, so pair sums are taken
    in a different order than in the single-context run and rounding differences grow with the steps: the bound here is
    1e-3 of max|v| / 1e-5 m (dp/4000) after 400 steps, not the 10-step tolerance (measured on B200: Symplectic 3e-7 m and
    4e-7, Verlet 1.2e-6 m and 1.2e-5 -- Verlet damps perturbations less)."""
    H.write_case(tmp_path, dp=0.04)
    case = str(tmp_path / "Case")
    options += " -cellmode:half"
    ref = dsb.Simulation(case, options=options)
    ss = slab.SlabSimulation(case, options=options, nranks=3, dist=None)
    nsteps = 400
    for s in range(nsteps):
        dt_ref, dt = ref.ctx.compute_step(), ss.step()
        assert abs(dt - dt_ref) <= 1e-3 * dt_ref, (s, dt, dt_ref)
    ids, pos, vr = ss.gather_byid()
    rid = ref.ctx.array("idp"); o = np.argsort(rid, kind="stable")
    assert np.array_equal(ids, rid[o]), "particles lost or duplicated by the migration"
    epos = float(np.max(np.abs(pos - ref.ctx.array("pos")[o])))
    evr = H.rel_err(vr, ref.ctx.array("velrho")[o])
    print(f"\n[slab migration {options}] migrated={ss.migrated_total()} owned={ss.owned_counts()} pos_err={epos:.3e} velrho_err={evr:.3e}")
    assert ss.migrated_total() > 0, "the case did not exercise migration"
    assert epos < 1e-5 and evr < 1e-3
    ss.close(); ref.close()


def test_slab_recut(tmp_path, built_lib):
    """Re-cutting between steps: forced cuts that move a cut by several planes (particles travel across a whole slab in
    two hops), then balanced cuts from the all-reduced histogram; the union of the slabs must stay the particle set of the
    single-context run and the state must agree as in the migration test."""
    H.write_case(tmp_path, dp=0.04)
    case = str(tmp_path / "Case")
    options = "-symplectic -cellmode:half"
    ref = dsb.Simulation(case, options=options)
    ss = slab.SlabSimulation(case, options=options, nranks=3, dist=None)
    c0 = list(ss.cuts)
    n0 = ref.ctx.info().np
    for s in range(30):
        ref.ctx.compute_step(); ss.step()
        if s == 9:    # squeeze the middle slab to two planes on the far side: everything it owned moves two slabs over
            forced = [c0[0], c0[2] + 2, c0[2] + 4, c0[3]]
            assert ss.recut(cuts=forced) == forced
            assert sum(ss.owned_counts()) == n0
        if s == 19:
            balanced = ss.recut()
            # the cuts balance WORK (a fluid particle weighs 3 boundary particles, csrc/slab_cuts.h), not heads
            work = []
            for r in ss.ranks:
                code = r.ctx.array("code"); own = (code & 0xe000) == 0
                work.append(3.0 * np.sum(own & ((code & 0x1800) >= 0x1000)) + np.sum(own & ((code & 0x1800) < 0x1000)))
            work = np.array(work)
            assert sum(ss.owned_counts()) == n0 and work.max() / work.mean() < 1.5, (balanced, work, ss.owned_counts())
    ids, pos, vr = ss.gather_byid()
    rid = ref.ctx.array("idp"); o = np.argsort(rid, kind="stable")
    assert np.array_equal(ids, rid[o])
    epos = float(np.max(np.abs(pos - ref.ctx.array("pos")[o])))
    evr = H.rel_err(vr, ref.ctx.array("velrho")[o])
    print(f"\n[slab recut] cuts {c0} -> {ss.cuts} migrated={ss.migrated_total()} owned={ss.owned_counts()} pos_err={epos:.3e} velrho_err={evr:.3e}")
    assert ss.migrated_total() > 0 and epos < 1e-6 and evr < 1e-4
    ss.close(); ref.close()


@pytest.mark.parametrize("nranks", [2, 3])
def test_slab_floating_body(nranks, tmp_path, built_lib):
    """A floating box that straddles a cut: every rank sums the force/torque of the floating particles it owns, the sums are
    all-reduced and every rank integrates the same rigid body.  Against the single-context run: body centre 1e-8 m, linear /
    angular velocity 1e-5 relative (the partial sums are added in a different order), particles as in the other slab tests."""
    H.write_case(tmp_path, dp=0.05, floating_box=(0.15, 0.25, 0.2, 0.1, 0.15, 0.1), floating_rel_density=0.5)
    case = str(tmp_path / "Case")
    ref = dsb.Simulation(case, options="-symplectic -cellmode:half")
    ss = slab.SlabSimulation(case, options="-symplectic -cellmode:half", nranks=nranks, dist=None)
    # put a cut through the middle of the box (x=0.2): the body must really be shared by at least two ranks
    p = ss.ranks[0].p
    cmid = int((0.2 - p.map_posmin[0]) / float(p.scell))
    ncx = int(p.domcells[0])
    ss.recut(cuts=[0, cmid, ncx] if nranks == 2 else [0, cmid, cmid + 2, ncx])   # slabs are >= 2 planes (halo width)
    owners = 0
    for r in ss.ranks:
        code = r.ctx.array("code")
        owners += int(np.any(((code & 0xe000) == 0) & ((code & 0x1800) == 0x1000)))
    assert owners >= 2, (ss.cuts, owners)
    for s in range(20):
        dt_ref, dt = ref.ctx.compute_step(), ss.step()
        assert abs(dt - dt_ref) <= 2e-5 * dt_ref, (s, dt, dt_ref)
    f0 = ref.ctx.get_floatings(1)[0]
    for r in ss.ranks:
        f = r.ctx.get_floatings(1)[0]
        assert max(abs(f.center[k] - f0.center[k]) for k in range(3)) < 1e-8
        v0 = np.array([*f0.fvel, *f0.fomega]); v = np.array([*f.fvel, *f.fomega])
        assert np.max(np.abs(v - v0)) < 1e-5 * max(np.max(np.abs(v0)), 1e-2), (v, v0)
    ids, pos, vr = ss.gather_byid()
    rid = ref.ctx.array("idp"); o = np.argsort(rid, kind="stable")
    assert np.array_equal(ids, rid[o])
    epos = float(np.max(np.abs(pos - ref.ctx.array("pos")[o])))
    evr = H.rel_err(vr, ref.ctx.array("velrho")[o])
    print(f"\n[slab floating x{nranks}] cuts={ss.cuts} owners={owners} pos_err={epos:.3e} velrho_err={evr:.3e}")
    assert epos < 1e-7 and evr < 2e-6
    ss.close(); ref.close()


def test_slab_piston_and_floating(tmp_path, built_lib):
    """BASELINE configs[3] in miniature: piston wavemaker (mvrectsinu) + floating box, 3 slabs, against the single context."""
    H.write_case(tmp_path, dp=0.05, piston=True, piston_ampl=0.05, piston_freq=2.0, floating_box=(0.45, 0.25, 0.2, 0.1, 0.15, 0.1), floating_rel_density=0.5)
    case = str(tmp_path / "Case")
    opts = "-symplectic -cellmode:half"
    ref = dsb.Simulation(case, options=opts)
    ss = slab.SlabSimulation(case, options=opts, nranks=3, dist=None)
    for s in range(40):
        dt_ref, dt = ref.ctx.compute_step(), ss.step()
        assert abs(dt - dt_ref) <= 2e-5 * dt_ref, (s, dt, dt_ref)
    f0 = ref.ctx.get_floatings(1)[0]
    f = ss.ranks[0].ctx.get_floatings(1)[0]
    assert max(abs(f.center[k] - f0.center[k]) for k in range(3)) < 1e-8
    ids, pos, vr = ss.gather_byid()
    rid = ref.ctx.array("idp"); o = np.argsort(rid, kind="stable")
    assert np.array_equal(ids, rid[o])
    epos = float(np.max(np.abs(pos - ref.ctx.array("pos")[o])))
    evr = H.rel_err(vr, ref.ctx.array("velrho")[o])
    # the piston must have moved (moving boundary particles are displaced identically on every rank that owns some)
    moving = (ref.ctx.array("code")[o] & 0x1800) == 0x0800
    assert moving.any() and np.max(np.abs(pos[moving, 0] - ref.ctx.array("pos")[o][moving, 0])) < 1e-12
    print(f"\n[slab piston+floating] cuts={ss.cuts} pos_err={epos:.3e} velrho_err={evr:.3e}")
    assert epos < 1e-7 and evr < 2e-6
    ss.close(); ref.close()


def test_slab_too_thin_is_loud(tmp_path, built_lib):
    H.write_case(tmp_path, dp=0.04)
    ss = slab.SlabSimulation(str(tmp_path / "Case"), options="-cellmode:half", nranks=2, dist=None)
    with pytest.raises(dsb.B2SError):
        ss.recut(cuts=[0, 1, ss.cuts[-1]])      # one plane < halo width of two planes
    ss.close()


@pytest.mark.parametrize("slipmode", [1, 2])
@pytest.mark.parametrize("nranks", [2, 3])
def test_slab_mdbc_lamsps(nranks, slipmode, tmp_path, built_lib):
    """BASELINE configs[2] in miniature: mDBC boundaries + laminar+SPS viscosity over slabs.  The halo is two kernel radii wide
    so that every rank can recompute the ghost-node density of its boundary ghosts; those sums visit the same fluid particles
    in another order than on the owner, so the result equals the single-context run to rounding (not bit for bit)."""
    H.write_case(tmp_path, dp=0.04, layers=4, normals=True,
                 params={"Boundary": 2, "SlipMode": slipmode, "ViscoTreatment": 2, "Visco": 1e-6, "DensityDT": 3})
    case = str(tmp_path / "Case")
    opts = "-cellmode:half"
    ref = dsb.Simulation(case, options=opts)
    ss = slab.SlabSimulation(case, options=opts, nranks=nranks, dist=None)
    assert ss.ranks[0].halo_planes == 4
    for s in range(10):
        dt_ref, dt = ref.ctx.compute_step(), ss.step()
        assert abs(dt - dt_ref) <= 2e-5 * dt_ref, (s, dt, dt_ref)
    ids, pos, vr = ss.gather_byid()
    rid = ref.ctx.array("idp"); o = np.argsort(rid, kind="stable")
    assert np.array_equal(ids, rid[o])
    epos = float(np.max(np.abs(pos - ref.ctx.array("pos")[o])))
    evr = H.rel_err(vr, ref.ctx.array("velrho")[o])
    print(f"\n[slab mdbc(slip {slipmode})+lamsps x{nranks}] cuts={ss.cuts} owned={ss.owned_counts()} pos_err={epos:.3e} velrho_err={evr:.3e}")
    assert epos < 1e-7 and evr < 1e-5      # the mDBC tolerance of the golden-fixture tests (ill-conditioned 4x4 inverse)
    ss.close(); ref.close()


@pytest.mark.parametrize("nranks", [2, 3])
def test_slab_periodic_y(nranks, tmp_path, built_lib):
    """Periodic boundaries across the slab axis' normal (Y periodic, slabs along x): every rank duplicates its owned particles AND
    the halo ghosts it received next to the periodic faces (corner coverage), so the slabs reproduce the single-context run.
    The fluid moves at 3 m/s along y, so particles wrap around during the 30 steps."""
    H.write_case(tmp_path, dp=0.03, open_axes="y", fluid_vel=(0.0, 3.0, 0.0), params={"YPeriodicIncZ": 0})
    case = str(tmp_path / "Case")
    ref = dsb.Simulation(case)
    ss = slab.SlabSimulation(case, nranks=nranks, dist=None)
    for s in range(30):
        dt_ref = ref.ctx.compute_step()
        dt = ss.step()
        assert abs(dt - dt_ref) <= 2e-5 * dt_ref, (s, dt, dt_ref)
    ids, pos, vr = ss.gather_byid()
    own = (ref.ctx.array("code") & 0xe000) == 0          # without the periodic duplicates
    rid = ref.ctx.array("idp")[own]; o = np.argsort(rid, kind="stable")
    rpos, rvr = ref.ctx.array("pos")[own][o], ref.ctx.array("velrho")[own][o]
    assert np.array_equal(ids, rid[o]), "particle set differs from the single-context run"
    epos, evr = float(np.max(np.abs(pos - rpos))), H.rel_err(vr, rvr)
    y0 = ref.params.maprealposmin[1]; ly = ref.params.maprealposmax[1] - y0
    assert np.all((rpos[:, 1] >= y0) & (rpos[:, 1] < y0 + ly)), "particles left the periodic range"
    print(f"\n[slab periodic-y x{nranks}] cuts={ss.cuts} pos_err={epos:.3e} velrho_err={evr:.3e}")
    assert epos < 1e-7 and evr < 2e-6
    ss.close(); ref.close()


def test_slab_gauges(tmp_path, built_lib):
    """Gauges in slab runs (csrc/slab_comm.cu grp_gauges): velocity points, an SWL line, a force gauge on the piston and a maxz gauge,
    three slabs against the single context.  The rank whose slab holds a point evaluates it, the force gauge sums over the owners of
    the boundary particles, maxz takes the maximum; every rank ends with the same numbers.  The single context evaluates before the
    end-of-step divide (the reference's order), the slabs after it: same particles, the sums in another order -> 2e-5 of the scale."""
    H.write_case(tmp_path, dp=0.025, piston=True, piston_ampl=0.05, piston_freq=2.0)
    case = str(tmp_path / "Case")
    opts = "-symplectic -cellmode:half"
    ref = dsb.Simulation(case, options=opts)
    ss = slab.SlabSimulation(case, options=opts, nranks=3, dist=None)
    assert ss.native, ss.cuts                                        # the library's step (the gauges are part of it) needs slabs wider than the halo
    p = ref.params
    code, idp, pos = ref.ctx.array("code"), ref.ctx.array("idp"), ref.ctx.array("pos")
    mov = np.nonzero((code & 0x1800) == 0x0800)[0]
    assert mov.size
    fluid = pos[(code & 0x1800) == 0x1800]
    scell, dom0 = float(p.scell), float(p.domposmin[0])
    xcuts = [dom0 + c * scell for c in ss.cuts[1:-1]]                 # x of the two inner cut planes
    ymid, zlow = float(fluid[:, 1].mean()), float(fluid[:, 2].min()) + 3 * float(p.dp)
    points = [(x + dx, ymid, zlow) for x in xcuts for dx in (-0.3 * scell, 0.3 * scell)] + [(float(fluid[:, 0].mean()), ymid, zlow)]
    ctxs = [ref.ctx] + [r.ctx for r in ss.ranks]
    gids = []
    for c in ctxs:
        g = [c.gauge_add_velocity(pt) for pt in points]
        g.append(c.gauge_add_swl((xcuts[0], ymid, float(fluid[:, 2].min())), (xcuts[0], ymid, float(fluid[:, 2].max()) + 0.1), float(p.dp), 0.5 * float(p.massfluid)))
        g.append(c.gauge_add_force(int(idp[mov].min()), int(mov.size), int(code[mov[0]] & 0x1fff)))
        g.append(c.gauge_add_maxz((xcuts[-1], ymid, float(fluid[:, 2].min())), 1.0, 4 * float(p.dp)))
        gids.append(g)
    assert all(g == gids[0] for g in gids)
    ref.ctx.gauges_compute(); ss.ranks[0].ctx.slab_gauges_compute()   # the measurement before the first step

    kforce = len(gids[0]) - 2

    def compare(label, skip=()):
        worst = 0.0
        for k, gid in enumerate(gids[0]):
            if k in skip:
                continue
            r0 = np.array(ref.ctx.gauge_result(gid))
            rs = [np.array(r.ctx.gauge_result(gid)) for r in ss.ranks]
            assert all(np.array_equal(rs[0], x) for x in rs), (label, k, rs)       # every rank holds the same results
            scale = max(float(np.abs(r0).max()), 1e-3)
            worst = max(worst, float(np.abs(rs[0] - r0).max()) / scale)
            assert np.abs(rs[0] - r0).max() <= 2e-5 * scale, (label, k, r0, rs[0])
        return worst
    w0 = compare("initial")
    for s in range(30):
        dt_ref, dt = ref.ctx.compute_step(), ss.step()
        assert abs(dt - dt_ref) <= 2e-5 * dt_ref, (s, dt, dt_ref)
    # In its step the single context evaluates before RunMotion and the end-of-step divide (the reference's order,
    # JSphCpuSingle.cpp:1192), the slabs after them: identical for everything but a force gauge on a block that moves -- the piston
    # is one step's displacement further for the slabs.  Evaluated on the same state (b2s_gauges_compute after the step) it agrees too.
    w1 = compare("after 30 steps, in-step values", skip=(kforce,))
    ref.ctx.gauges_compute()
    w2 = compare("after 30 steps, same state")
    w1 = max(w1, w2)
    fz = np.array(ref.ctx.gauge_result(gids[0][-2]))
    vel = np.array(ref.ctx.gauge_result(gids[0][0]))
    print(f"\n[slab gauges] cuts={ss.cuts} worst rel diff: initial {w0:.2e}, after 30 steps {w1:.2e}; piston force {fz}, velocity at the first point {vel}")
    assert np.abs(fz).max() > 0 and np.abs(vel).max() > 0            # the gauges measure something
    ss.close(); ref.close()
