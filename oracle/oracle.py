"""TEST INFRASTRUCTURE ONLY -- Python driver of the C restatement (oracle/sph_oracle.c).

Never imported by the product package; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg use it.  ``OracleSim`` strings the C functions together in the order of JSphCpuSingle::ComputeStep_Sym /
ComputeStep_Ver + the Run-loop tail (src/source/JSphCpuSingle.cpp:871-932,1188-1196) so that whole steps can
be compared with the CUDA path and with the dumps of the compiled reference (oracle/_ref/dsph_ref_probe).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "_build", "libsph_oracle.so")
REF_DIR = os.path.join(_HERE, "_ref")
REF_PROBE = os.path.join(REF_DIR, "dsph_ref_probe")
REF_SOLVER = os.path.join(REF_DIR, "DualSPHysics5.4CPU_linux64")


class Cte(C.Structure):
    _fields_ = [
        ("tkernel", C.c_int32), ("tvisco", C.c_int32), ("tdensity", C.c_int32), ("tboundary", C.c_int32),
        ("tstep", C.c_int32), ("use_floating", C.c_int32), ("dt_allparticles", C.c_int32), ("scelldiv", C.c_int32),
        ("kernelh", C.c_float), ("kernelsize2", C.c_float), ("awen", C.c_float), ("bwenh", C.c_float),
        ("cub_a2", C.c_float), ("cub_a24", C.c_float), ("cub_c1", C.c_float), ("cub_d1", C.c_float),
        ("cub_c2", C.c_float), ("cub_od_wdeltap", C.c_float),
        ("cteb", C.c_float), ("gamma", C.c_float), ("rhopzero", C.c_float), ("eta2", C.c_float),
        ("massfluid", C.c_float), ("massbound", C.c_float),
        ("ddtkh", C.c_float), ("ddtgz", C.c_float), ("visco", C.c_float), ("viscoboundfactor", C.c_float),
        ("spssmag", C.c_float), ("spsblin", C.c_float),
        ("scell", C.c_float), ("movlimit", C.c_float), ("rhopoutmin", C.c_float), ("rhopoutmax", C.c_float),
        ("gravity", C.c_float * 3),
        ("cs0", C.c_double), ("cflnumber", C.c_double), ("dtmin", C.c_double),
        ("domposmin", C.c_double * 3), ("maprealposmin", C.c_double * 3), ("maprealsize", C.c_double * 3),
        ("domcellcode", C.c_uint32), ("domcells", C.c_uint32 * 3),
        ("shiftmode", C.c_int32), ("shiftcoef", C.c_float), ("shifttfs", C.c_float), ("dp", C.c_float),
        ("periactive", C.c_int32), ("simulate2d", C.c_int32), ("perinc", (C.c_double * 3) * 3),
        ("mapposmin", C.c_double * 3), ("mapposmax", C.c_double * 3),
    ]


class Div(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "cellmin_x", "cellmin_y", "cellmin_z", "cellmax_x", "cellmax_y", "cellmax_z", "ncx", "ncy", "ncz", "nsheet",
        "nct", "nctt", "box_boundignore", "box_fluid", "box_boundout", "box_fluidout", "box_boundoutignore",
        "box_fluidoutignore", "np", "npb", "npbok", "npb_ignore", "npb_out", "npf_out")]


def build(force: bool = False) -> str:
    """gcc -O2 -fopenmp -shared oracle/sph_oracle.c -> oracle/_build/libsph_oracle.so"""
    src = os.path.join(_HERE, "sph_oracle.c")
    if force or not os.path.exists(SO_PATH) or os.path.getmtime(SO_PATH) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(SO_PATH), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-o", SO_PATH, src, "-lm"])
    return SO_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(SO_PATH)
        _lib.orc_dtvariable.restype = C.c_double
        _lib.orc_dtvariable.argtypes = [C.POINTER(Cte), C.c_double, C.c_double, C.c_float]
        _lib.orc_sizeof_cte.restype = C.c_size_t
        _lib.orc_sizeof_div.restype = C.c_size_t
        assert _lib.orc_sizeof_cte() == C.sizeof(Cte) and _lib.orc_sizeof_div() == C.sizeof(Div)
    return _lib


def cte_from_params(p) -> Cte:
    """Builds the oracle constants from a (finalized) b2s Params-like object (same field names)."""
    c = Cte()
    c.tkernel, c.tvisco, c.tdensity, c.tboundary, c.tstep = p.tkernel, p.tvisco, p.tdensity, p.tboundary, p.tstep
    c.use_floating, c.dt_allparticles, c.scelldiv = p.use_floating, p.dt_allparticles, p.scelldiv
    c.kernelh, c.kernelsize2, c.awen, c.bwenh = p.kernelh, p.kernelsize2, p.awen, p.bwenh
    c.cub_a2, c.cub_a24, c.cub_c1, c.cub_d1 = p.cubic_a2, p.cubic_a24, p.cubic_c1, p.cubic_d1
    c.cub_c2, c.cub_od_wdeltap = p.cubic_c2, p.cubic_od_wdeltap
    c.cteb, c.gamma, c.rhopzero, c.eta2 = p.cteb, p.gamma, p.rhopzero, p.eta2
    c.massfluid, c.massbound = p.massfluid, p.massbound
    c.ddtkh, c.ddtgz, c.visco, c.viscoboundfactor = p.ddtkh, p.ddtgz, p.visco, p.viscoboundfactor
    c.spssmag, c.spsblin = p.spssmag, p.spsblin
    c.scell, c.movlimit, c.rhopoutmin, c.rhopoutmax = p.scell, p.movlimit, p.rhopoutmin, p.rhopoutmax
    for k in range(3):
        c.gravity[k] = p.gravity[k]
        c.domposmin[k] = p.domposmin[k]
        c.maprealposmin[k] = p.maprealposmin[k]
        c.maprealsize[k] = p.maprealsize[k]
        c.domcells[k] = p.domcells[k]
    c.cs0, c.cflnumber, c.dtmin = p.cs0, p.cflnumber, p.dtmin
    c.domcellcode = p.domcellcode
    c.shiftmode, c.shiftcoef, c.shifttfs, c.dp = p.shiftmode, p.shiftcoef, p.shifttfs, p.dp
    c.periactive = p.periactive
    c.simulate2d = int(getattr(p, "simulate2d", 0))
    for k in range(3):
        c.perinc[0][k], c.perinc[1][k], c.perinc[2][k] = p.perixinc[k], p.periyinc[k], p.perizinc[k]
        c.mapposmin[k], c.mapposmax[k] = p.map_posmin[k], p.map_posmax[k]
    return c


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleSim:
    """Whole-step CPU oracle over numpy arrays (cell-sorted order, like the solver's state)."""

    def __init__(self, cte: Cte, idp, code, pos, velrho, npb: int, dtini: float, verletsteps: int = 40,
                 ftomassp=None, boundnor=None, mdbc_corrector=False, damping=None, motion=None, floatings=None, ftpause=0.0,
                 slipmode=1):
        self.L = lib()
        self.c = cte
        self.idp = np.ascontiguousarray(idp, np.uint32).copy()
        self.code = np.ascontiguousarray(code, np.uint16).copy()
        self.pos = np.ascontiguousarray(pos, np.float64).reshape(-1, 3).copy()
        self.velrho = np.ascontiguousarray(velrho, np.float32).reshape(-1, 4).copy()
        self.np_, self.npb = int(self.idp.shape[0]), int(npb)
        self.dcell = np.zeros(self.np_, np.uint32)
        self.L.orc_load_dcell(C.byref(self.c), self.np_, _p(self.code), _p(self.pos), _p(self.dcell))
        self.tstep = cte.tstep
        self.velrhom1 = self.velrho.copy() if self.tstep == 1 else None
        self.tau = np.zeros((self.np_, 6), np.float32) if cte.tvisco == 2 else None
        self.ftomassp = None if ftomassp is None else np.ascontiguousarray(ftomassp, np.float32)
        self.boundnor = None
        if cte.tboundary == 2:
            self.boundnor = np.zeros((self.np_, 3), np.float32)
            if boundnor is not None:
                bn = np.asarray(boundnor, np.float32).reshape(-1, 3)
                self.boundnor[:bn.shape[0]] = bn
        self.gauges = None                            # GaugeOracle, attached by the caller
        self.slipmode = int(slipmode)                 # 1 vel0, 2 no-slip, 3 free-slip (mDBC2)
        self.mdbc_corrector = mdbc_corrector or self.slipmode >= 2      # MdbcCorrector (JSph.cpp:899)
        self.boundmode = self.tangenvel = None        # mDBC2: per interaction, in the current sorted order
        self.motionace = np.zeros((self.np_, 3), np.float32) if (self.boundnor is not None and self.slipmode >= 2) else None
        # damping zones: (nzones, 64) doubles in the layout of orc_damping (built by damping_zones() below)
        self.damping = None if damping is None or len(damping) == 0 else np.ascontiguousarray(damping, np.float64).reshape(-1, 64)
        # predefined motion: (MotionOracle, idbegin, count) of the moving block, or None
        self.motion = motion
        # floating bodies: list of dicts (begin, count, mass, massp, center[3] f64, angles/fvel/fomega[3] f32, inertiaini[3,3] f32)
        self.floatings, self.ftpause = floatings or [], float(ftpause)
        self.symplectic_dtpre = float(dtini)
        self.timestep, self.nstep, self.verletstep, self.verletsteps = 0.0, 0, 0, verletsteps
        self.div = Div()
        self.npf_out_total = 0
        nctt = 2 * int(cte.domcells[0]) * int(cte.domcells[1]) * int(cte.domcells[2]) + 8
        self.begincell = np.zeros(nctt, np.uint32)
        self.cellpart = np.zeros(self.np_, np.uint32)
        self.sortpart = np.zeros(self.np_, np.uint32)
        self.pospre = self.velrhopre = None
        self.celldiv()

    _PER_ARRAYS = ("idp", "code", "dcell", "pos", "velrho", "velrhom1", "tau", "pospre", "velrhopre", "boundnor", "motionace")

    def periodic(self):
        """JSphCpuSingle::RunPeriodic (JSphCpuSingle.cpp:416-504) with PeriodicMakeList / PeriodicDuplicate* (:210-372): the old
        duplicates are marked to be ignored, new ones are appended behind the particles -- boundary then fluid; per type the axes
        X, Y, Z; per axis first from the duplicates of this type made so far, then from the original particles; per particle the
        +inc copy before the -inc copy.  That order decides the particle order inside the cells after the stable sort."""
        c = self.c
        n0, npb0 = self.np_, self.npb
        for name in self._PER_ARRAYS:
            a = getattr(self, name)
            if a is not None:
                setattr(self, name, a[:n0])
        per = (self.code & 0xe000) == 0x2000
        self.code[per] = (self.code[per] & 0x1fff) | 0x4000          # PeriodicIgnore
        mapmin = np.array([c.mapposmin[k] for k in range(3)]); mapmax = np.array([c.mapposmax[k] for k in range(3)])
        dommin = np.array([c.domposmin[k] for k in range(3)]); domcells = np.array([c.domcells[k] for k in range(3)], np.int64)
        scell = np.float64(np.float32(c.scell))
        dcc = int(c.domcellcode)
        shx, shy = (dcc >> 10) & 31, (dcc >> 15) & 31
        nper = [0, 0]
        for ctype in (0, 1):
            for ax in range(3):
                if not (c.periactive >> ax) & 1:
                    continue
                inc = np.array([c.perinc[ax][k] for k in range(3)])
                for cblock in (0, 1):
                    ntot = self.idp.shape[0]
                    idx = (np.arange(0, npb0) if ctype == 0 else np.arange(npb0, n0)) if cblock else np.arange(ntot - nper[ctype], ntot)
                    if idx.size == 0:
                        continue
                    idx = idx[(self.code[idx] & 0xe000) <= 0x2000]
                    ps = self.pos[idx]
                    plus = np.all((mapmin <= ps + inc) & (ps + inc < mapmax), axis=1)
                    minus = np.all((mapmin <= ps - inc) & (ps - inc < mapmax), axis=1)
                    src = np.concatenate([idx[plus], idx[minus]])
                    sign = np.concatenate([np.ones(int(plus.sum())), -np.ones(int(minus.sum()))])
                    order = np.lexsort((-sign, src))                   # by particle, + before -
                    src, sign = src[order], sign[order]
                    if src.size == 0:
                        continue
                    newpos = self.pos[src] + sign[:, None] * inc
                    cell = ((newpos - dommin) / scell).astype(np.uint32).astype(np.int64)   # unsigned((ps-DomPosMin)/Scell)
                    cell = np.minimum(cell, domcells)                                     # PeriodicDuplicatePos: clamped to cellmax=DomCells
                    newdcell = ((cell[:, 0] << shx) | (cell[:, 1] << shy) | cell[:, 2]).astype(np.uint32)
                    for name in self._PER_ARRAYS:
                        a = getattr(self, name)
                        if a is None:
                            continue
                        if name == "pos":
                            add = newpos
                        elif name == "dcell":
                            add = newdcell
                        elif name == "code":
                            add = ((a[src] & 0x1fff) | 0x2000).astype(np.uint16)
                        else:
                            add = a[src]
                        setattr(self, name, np.ascontiguousarray(np.concatenate([a, add.astype(a.dtype)])))
                    nper[ctype] += int(src.size)
        self.np_ = int(self.idp.shape[0])
        self.npb = npb0 + nper[0]      # bound-type particles entering the divide (the sort groups them by type)
        if self.cellpart.shape[0] < self.np_:
            self.cellpart = np.zeros(self.np_ + 1024, np.uint32)
            self.sortpart = np.zeros(self.np_ + 1024, np.uint32)

    # RunCellDivide (JSphCpuSingle.cpp:497-586)
    def celldiv(self):
        if self.c.periactive:
            self.periodic()
        n = self.np_
        r = self.L.orc_celldiv(C.byref(self.c), n, self.npb, _p(self.dcell), _p(self.code), _p(self.cellpart),
                               _p(self.sortpart), _p(self.begincell), C.byref(self.div))
        assert r == 0
        sp = self.sortpart[:n]
        for name in self._PER_ARRAYS:
            a = getattr(self, name)
            if a is not None:
                setattr(self, name, np.ascontiguousarray(a[:n][sp][:int(self.div.np)]))   # excluded / ignored particles sort to the end: dropped
        self.npf_out_total += int(self.div.npf_out)
        self.np_, self.npb = int(self.div.np), int(self.div.npb)
        if self.div.npb_out:
            raise RuntimeError("oracle: boundary particles excluded")

    def mdbc(self):
        """MdbcBoundCorrection (JSphCpuSingle.cpp:745-760), SlipMode vel0."""
        if self.boundnor is None:
            return
        if self.slipmode >= 2:     # Interaction_Mdbc2Correction over the Npb boundary particles (JSphCpuSingle.cpp:763-771)
            nb = self.npb
            self.boundmode = np.zeros(max(self.np_, 1), np.uint8)
            self.tangenvel = np.zeros((max(self.np_, 1), 3), np.float32)
            self.L.orc_mdbc2(C.byref(self.c), C.byref(self.div), nb, _p(self.begincell), _p(self.pos), _p(self.code), _p(self.boundnor),
                             _p(self.motionace), 1 if self.slipmode == 3 else 0, _p(self.velrho), _p(self.boundmode), _p(self.tangenvel))
            return
        n = self.np_ if self.c.use_floating else int(self.div.npbok)
        self.L.orc_mdbc_vel0(C.byref(self.c), C.byref(self.div), n, _p(self.begincell), _p(self.pos), _p(self.code),
                             _p(self.boundnor), _p(self.velrho))

    def interaction(self, corrector=False):
        n = self.np_
        if not corrector or self.mdbc_corrector:
            self.mdbc()
        self.ace = np.zeros((n, 3), np.float32)
        self.ar = np.zeros(n, np.float32)
        self.press = np.zeros(n, np.float32)
        vm, am, vd = C.c_double(0), C.c_double(0), C.c_float(0)
        self.L.orc_interaction(C.byref(self.c), C.byref(self.div), _p(self.begincell), _p(self.dcell), _p(self.pos),
                               _p(self.velrho), _p(self.code), None if self.tau is None else _p(self.tau),
                               None if self.ftomassp is None else _p(self.ftomassp),
                               _p(self.ace), _p(self.ar), _p(self.press), C.byref(vm), C.byref(am), C.byref(vd),
                               None if self.boundmode is None else _p(self.boundmode), None if self.tangenvel is None else _p(self.tangenvel))
        self.velmax, self.acemax, self.viscdtmax = vm.value, am.value, vd.value
        return self.L.orc_dtvariable(C.byref(self.c), vm.value, am.value, vd.value)

    def shifting(self, dt):
        """RunShifting(dt) (JSphCpuSingle.cpp:881,924): displacement of the coming position update, or None."""
        if not self.c.shiftmode:
            return None
        sh = np.zeros((self.np_, 4), np.float32)
        self.L.orc_shifting(C.byref(self.c), C.byref(self.div), _p(self.begincell), _p(self.dcell), _p(self.pos), _p(self.velrho),
                            _p(self.code), None if self.ftomassp is None else _p(self.ftomassp), C.c_double(dt), _p(sh))
        return sh

    def run_damping(self, dt):
        """RunDamping(dt) (JSphCpuSingle.cpp:885,928)."""
        if self.damping is not None:
            self.L.orc_damping(self.np_, self.npb, C.c_double(dt), int(self.damping.shape[0]), _p(self.damping), _p(self.pos),
                               _p(self.code), _p(self.velrho))

    def run_floating(self, dt, predictor):
        """JSphCpuSingle::RunFloating (JSphCpuSingle.cpp:938-987): FtPartsSumAce (:1013-1039), JSph::FtComputeAceVel
        (JSph.cpp:2565-2660, free bodies), FtPartsUpdate (:1044-1100), JSph::FtUpdateFloatings (JSph.cpp:2687-2701).
        Sums run over the body's particles in id order, in float like the reference."""
        if not self.floatings or self.timestep < self.ftpause:
            return
        f32 = np.float32
        n = self.np_
        grav = np.array([self.c.gravity[k] for k in range(3)], f32)
        for ft in self.floatings:
            sel = np.nonzero((self.idp[:n] >= ft["begin"]) & (self.idp[:n] < ft["begin"] + ft["count"]))[0]
            sel = sel[np.argsort(self.idp[sel], kind="stable")].astype(np.uint32)
            ace = self.ace[sel].astype(f32)
            dist = (self.pos[sel] - ft["center"]).astype(f32)
            acelin = np.array([np.cumsum(ace[:, k], dtype=f32)[-1] for k in range(3)], f32)
            t = np.stack([ace[:, 2] * dist[:, 1] - ace[:, 1] * dist[:, 2], ace[:, 0] * dist[:, 2] - ace[:, 2] * dist[:, 0],
                          ace[:, 1] * dist[:, 0] - ace[:, 0] * dist[:, 1]], axis=1).astype(f32)
            aceang = np.array([np.cumsum(t[:, k], dtype=f32)[-1] for k in range(3)], f32)
            fmassp, fmass = f32(ft["massp"]), f32(ft["mass"])
            flin, fang = (acelin * fmassp).astype(f32), (aceang * fmassp).astype(f32)
            alin = ((flin + grav * fmass) / fmass).astype(f32)
            a = ft["angles"].astype(f32)
            cx, cy, cz, sx, sy, sz = np.cos(a[0]), np.cos(a[1]), np.cos(a[2]), np.sin(a[0]), np.sin(a[1]), np.sin(a[2])
            rot = np.array([[cy * cz, -cy * sz, sy], [sx * sy * cz + cx * sz, -sx * sy * sz + cx * cz, -sx * cy],
                            [-cx * sy * cz + sx * sz, cx * sy * sz + sx * cz, cx * cy]], f32)
            inert = ((rot @ ft["inertiaini"].astype(f32)).astype(f32) @ rot.T).astype(f32)
            aang = (np.linalg.inv(inert.astype(np.float64)).astype(f32) @ fang).astype(f32)
            fcenter = ft["center"] + dt * ft["fvel"].astype(np.float64)
            fvel = (dt * alin.astype(np.float64) + ft["fvel"].astype(np.float64)).astype(f32)
            fomega = (dt * aang.astype(np.float64) + ft["fomega"].astype(np.float64)).astype(f32)
            # FtPartsUpdate: every particle moves with its own velocity, then takes the rigid-body velocity at its new position
            mov = np.ascontiguousarray(dt * self.velrho[sel, :3].astype(np.float64))
            self.L.orc_move_particles(C.byref(self.c), int(sel.size), _p(sel), _p(mov), _p(self.pos), _p(self.dcell), _p(self.code))
            d2 = (self.pos[sel] - fcenter).astype(f32)
            self.velrho[sel, 0] = fvel[0] + (fomega[1] * d2[:, 2] - fomega[2] * d2[:, 1])
            self.velrho[sel, 1] = fvel[1] + (fomega[2] * d2[:, 0] - fomega[0] * d2[:, 2])
            self.velrho[sel, 2] = fvel[2] + (fomega[0] * d2[:, 1] - fomega[1] * d2[:, 0])
            if not predictor:
                ft["center"] = fcenter
                ft["angles"] = (ft["angles"].astype(np.float64) + fomega.astype(np.float64) * dt).astype(f32)
                ft["fvel"], ft["fomega"] = fvel, fomega

    def run_motion(self, dt):
        """CalcMotion + RunMotion (JSph.cpp CalcMotion, JSphCpu.cpp:2189-2210): MoveLinBound / MoveMatBound of the moving block."""
        if self.motion is None:
            return
        mo, idbegin, count = self.motion
        res = mo.proces_time(self.timestep, dt)
        n = self.np_
        sel = np.nonzero((self.idp[:n] >= idbegin) & (self.idp[:n] < idbegin + count) & ((self.code[:n] & 0xe000) == 0))[0].astype(np.uint32)
        if res is None:
            if self.motionace is not None:
                self.motionace[sel] = 0          # CopyMotionVelAce: the motion velocity did not change (JSphCpu_mdbc.cpp:472-490)
            return
        ps = self.pos[sel]
        if res[0] == "lin":
            mov = np.tile(res[1], (sel.size, 1))
        else:
            m = res[1]
            mov = (ps @ m[:3, :3].T + m[:3, 3]) - ps
            if self.boundnor is not None:
                gs = ps + self.boundnor[sel].astype(np.float64)
                gs2 = gs @ m[:3, :3].T + m[:3, 3]
                self.boundnor[sel] = (gs2 - (ps + mov)).astype(np.float32)
        mov = np.ascontiguousarray(mov, np.float64)
        self.L.orc_move_particles(C.byref(self.c), int(sel.size), _p(sel), _p(mov), _p(self.pos), _p(self.dcell), _p(self.code))
        vnew = (mov / dt).astype(np.float32)
        if self.motionace is not None:
            self.motionace[sel] = ((vnew.astype(np.float64) - self.velrho[sel, :3].astype(np.float64)) / dt).astype(np.float32)
        self.velrho[sel, :3] = vnew

    def step(self) -> float:
        n = self.np_
        if self.tstep == 2:
            dt = self.symplectic_dtpre
            dt_p = self.interaction()
            self.pospre, self.velrhopre = self.pos[:n].copy(), self.velrho[:n].copy()
            self.pos, self.velrho = self.pospre.copy(), self.velrhopre.copy()
            self.L.orc_symplectic_pre(C.byref(self.c), n, self.npb, C.c_double(dt), _p(self.pospre), _p(self.velrhopre),
                                      _p(self.ar), _p(self.ace), _p(self.pos), _p(self.velrho), _p(self.dcell), _p(self.code),
                                      None if self.boundmode is None else _p(self.boundmode))
            self.run_floating(dt * 0.5, True)
            self.celldiv()
            n = self.np_
            dt_c = self.interaction(corrector=True)
            sh = self.shifting(dt)
            self.L.orc_symplectic_corr(C.byref(self.c), n, self.npb, C.c_double(dt), _p(self.pospre), _p(self.velrhopre),
                                       _p(self.ar), _p(self.ace), _p(self.pos), _p(self.velrho), _p(self.dcell), _p(self.code),
                                       None if sh is None else _p(sh), None if self.boundmode is None else _p(self.boundmode))
            self.run_floating(dt, False)
            self.run_damping(dt)
            self.pospre = self.velrhopre = None
            self.symplectic_dtpre = min(dt_p, dt_c)
            self.dt_p, self.dt_c = dt_p, dt_c
        else:
            dt = self.interaction()
            self.verletstep += 1
            euler = not (self.verletstep < self.verletsteps)
            if euler:
                self.verletstep = 0
            v2 = self.velrho if euler else self.velrhom1
            new = np.zeros_like(self.velrho[:n])
            sh = self.shifting(dt)
            self.L.orc_verlet(C.byref(self.c), n, self.npb, C.c_double(dt), 1 if euler else 0, _p(self.velrho), _p(v2),
                              _p(self.ar), _p(self.ace), _p(self.pos), _p(self.dcell), _p(self.code), _p(new),
                              None if sh is None else _p(sh), None if self.boundmode is None else _p(self.boundmode))
            self.velrhom1 = self.velrho[:n]
            self.velrho = new
            self.run_floating(dt, False)
            self.run_damping(dt)
        if self.gauges is not None:
            self.gauges.update(self, self.timestep + dt)     # RunGaugeSystem(TimeStep+stepdt) (JSphCpuSingle.cpp:1192)
        self.run_motion(dt)
        self.celldiv()
        self.timestep += dt
        self.nstep += 1
        return dt


class GaugeOracle:
    """Velocity / SWL / MaxZ / force gauges restated over an OracleSim state (numpy): JGaugeVelocity, JGaugeSwl, JGaugeMaxZ,
    JGaugeForce::CalculeCpu (JDsGaugeItem.cpp:427-478, 728-790, 1023-1091, 1847-1908) with the timing rules of JGaugeItem
    (computedt / outputdt; JDsGaugeItem.h:274-275, .cpp:239-249).  Like the reference after ComputeStep, the neighbour search
    walks the cell table of the LAST divide while the particles already carry their new positions."""

    def __init__(self, xml_gauges, params, blocks, timemax, timeout):
        self.items, self.rows = [], {}
        df = {"cdt": timeout, "cs": 0.0, "ce": timemax, "odt": timeout, "os": 0.0, "oe": timemax, "out": False}

        def common(e, d):
            d = dict(d)
            if e is None:
                return d
            for tag, key, at in (("computedt", "cdt", "value"), ("outputdt", "odt", "value")):
                if e.find(tag) is not None:
                    d[key] = float(e.find(tag).get(at))
            for tag, ks, ke in (("computetime", "cs", "ce"), ("outputtime", "os", "oe")):
                c = e.find(tag)
                if c is not None:
                    if c.get("start") is not None:
                        d[ks] = float(c.get("start"))
                    if c.get("end") is not None:
                        d[ke] = float(c.get("end"))
            if e.find("output") is not None:
                d["out"] = e.find("output").get("value", "false").lower() in ("true", "1")
            return d
        df = common(xml_gauges.find("default"), df)
        f3 = lambda e: np.array([float(e.get(a)) for a in "xyz"])  # noqa: E731
        for e in xml_gauges:
            if e.tag in ("default",) or e.tag.startswith("_"):
                continue
            g = common(e, df)
            g.update({"type": e.tag, "name": e.get("name"), "cnext": 0.0, "onext": 0.0})
            if e.tag == "velocity":
                g["point"] = f3(e.find("point"))
            elif e.tag == "swl":
                g["p0"], g["p2"] = f3(e.find("point0")), f3(e.find("point2"))
                pd = e.find("pointdp")
                pointdp = params.dp * float(pd.get("coefdp")) if pd.get("coefdp") is not None else float(np.float32(pd.get("value")))
                ml = e.find("masslimit")
                g["masslimit"] = np.float32(ml.get("value")) if ml is not None and ml.get("value") is not None else (
                    np.float32(params.massfluid) * np.float32(ml.get("coef")) if ml is not None else np.float32(params.massfluid) * np.float32(0.5))
                dv = g["p2"] - g["p0"]
                dis = np.sqrt((dv * dv).sum())
                npt = int(dis / pointdp)                       # JGaugeSwl::SetPoints (JDsGaugeItem.cpp:614-633)
                if dis - pointdp * npt >= pointdp * 0.1:
                    npt += 1
                npt = max(npt, 1)
                g["np"], g["dir"] = npt, dv / dis * (dis / npt)
            elif e.tag == "maxz":
                g["p0"] = f3(e.find("point0"))
                g["height"] = float(np.float32(e.find("height").get("value")))
                dl = e.find("distlimit")
                g["distlimit"] = (np.float32(dl.get("value")) if dl.get("value") is not None else
                                  np.float32(params.dp * float(np.float32(dl.get("coefdp")))) if dl.get("coefdp") is not None else
                                  np.float32(float(params.kernelh) * float(np.float32(dl.get("coefh")))))
            elif e.tag == "force":
                mkb = int(e.find("target").get("mkbound"))
                idx = {"fixed": 0, "moving": 0, "floating": 0, "fluid": 0}
                tcode = {"fixed": 0x0000, "moving": 0x0800, "floating": 0x1000, "fluid": 0x1800}
                for tag, beg, n, node in blocks:
                    if tag != "fluid" and int(node.get("mkbound")) == mkb:
                        g["code"], g["begin"], g["count"] = tcode[tag] | idx[tag], beg, n
                    idx[tag] += 1
            else:
                raise ValueError(f"gauge <{e.tag}> is not restated")
            self.items.append(g)
            self.rows[g["name"]] = []

    @staticmethod
    def _wab(c, rr2):
        rad = np.sqrt(rr2); qq = rad / np.float32(c.kernelh)
        if c.tkernel == 2:
            w1 = np.float32(1) - np.float32(0.5) * qq
            return np.float32(c.awen) * (qq + qq + np.float32(1)) * (w1 * w1) * (w1 * w1)
        w = np.where(rad > np.float32(c.kernelh), np.float32(c.cub_a24) * (np.float32(2) - qq) ** 3,
                     np.float32(c.cub_a2) * (np.float32(1) + (np.float32(0.75) * qq - np.float32(1.5)) * qq * qq))
        return w.astype(np.float32)

    @staticmethod
    def _neigh(sim, pt, fluid=True):
        """Indices of the particles in the cell window of point pt (nsearch::Init(pos,...), JCellSearch_inline.h) in table order."""
        c, d = sim.c, sim.div
        sc = np.float64(np.float32(c.scell))
        cel = [int((pt[k] - c.domposmin[k]) / sc) - int((d.cellmin_x, d.cellmin_y, d.cellmin_z)[k]) for k in range(3)]
        n = (int(d.ncx), int(d.ncy), int(d.ncz)); s = int(c.scelldiv)
        ini = [cel[k] - min(cel[k], s) for k in range(3)]
        fin = [cel[k] + min(n[k] - cel[k] - 1, s) + 1 for k in range(3)]
        if ini[0] < 0 or ini[1] < 0 or ini[2] < 0 or fin[0] <= ini[0]:
            return np.zeros(0, np.int64)
        beg = sim.begincell[int(d.box_fluid):] if fluid else sim.begincell
        out = []
        for z in range(ini[2], fin[2]):
            for y in range(ini[1], fin[1]):
                a, b = int(beg[ini[0] + y * n[0] + z * int(d.nsheet)]), int(beg[fin[0] + y * n[0] + z * int(d.nsheet)])
                if b > a:
                    out.append(np.arange(a, b))
        return np.concatenate(out) if out else np.zeros(0, np.int64)

    def _point_sums(self, sim, pt):
        idx = self._neigh(sim, pt)
        if idx.size == 0:
            return 0.0, np.zeros(3)
        dr = (pt - sim.pos[idx]).astype(np.float32)
        rr2 = dr[:, 0] * dr[:, 0] + dr[:, 1] * dr[:, 1] + dr[:, 2] * dr[:, 2]
        ok = (rr2 <= np.float32(sim.c.kernelsize2)) & (rr2 >= np.float32(1e-18)) & ((sim.code[idx] & 0x1800) == 0x1800)
        idx, rr2 = idx[ok], rr2[ok]
        w = (self._wab(sim.c, rr2) * (np.float32(sim.c.massfluid) / sim.velrho[idx, 3])).astype(np.float32)
        return float(w.astype(np.float64).sum()), (w[:, None].astype(np.float64) * sim.velrho[idx, :3]).sum(axis=0)

    def measure(self, sim, g):
        c = sim.c
        if g["type"] == "velocity":
            _, sv = self._point_sums(sim, g["point"])
            return [np.float32(v) for v in sv] + [np.float32(v) for v in g["point"]]
        if g["type"] == "swl":
            mpre, surf, pt = np.float32(0), None, g["p0"].copy()
            for cp in range(g["np"] + 1):
                sw, _ = self._point_sums(sim, pt)
                mass = np.float32(sw * float(np.float32(c.massfluid)))
                if mass > g["masslimit"]:
                    mpre = mass
                if mass < g["masslimit"] and mpre:
                    fxm1 = (g["masslimit"] - mpre) / (mass - mpre) - np.float32(1)
                    surf = pt + g["dir"] * float(fxm1)
                    break
                pt = pt + g["dir"]
            if surf is None:
                surf = g["p0"] + g["dir"] * (g["np"] if mpre else 0)
            return [np.float32(v) for v in surf] + [np.float32(v) for v in g["p0"]] + [np.float32(v) for v in g["p2"]]
        if g["type"] == "maxz":
            d = sim.div
            sc = np.float64(np.float32(c.scell))
            cmin = (int(d.cellmin_x), int(d.cellmin_y), int(d.cellmin_z)); n = (int(d.ncx), int(d.ncy), int(d.ncz))
            cel = [int((g["p0"][k] - c.domposmin[k]) / sc) - cmin[k] for k in range(3)]
            cz2 = int((g["p0"][2] + g["height"] - c.domposmin[2]) / sc) - cmin[2]
            ncel = int(np.ceil(float(g["distlimit"]) / sc))
            x0, x1 = max(cel[0] - ncel, 0), min(cel[0] + ncel + 1, n[0])
            y0, y1 = max(cel[1] - ncel, 0), min(cel[1] + ncel + 1, n[1])
            z0, z1 = max(cel[2] - ncel, 0), min(cz2 + ncel + 1, n[2])
            zmax = -np.inf
            beg = sim.begincell[int(d.box_fluid):]
            if x0 < x1:
                for z in range(z0, z1):
                    for y in range(y0, y1):
                        a, b = int(beg[x0 + y * n[0] + z * int(d.nsheet)]), int(beg[x1 + y * n[0] + z * int(d.nsheet)])
                        if b <= a:
                            continue
                        idx = np.arange(a, b)
                        dx = (g["p0"][0] - sim.pos[idx, 0]).astype(np.float32); dy = (g["p0"][1] - sim.pos[idx, 1]).astype(np.float32)
                        ok = (dx * dx + dy * dy <= g["distlimit"] * g["distlimit"]) & ((sim.code[idx] & 0x1800) == 0x1800)
                        if ok.any():
                            zmax = max(zmax, float(np.float32(sim.pos[idx[ok], 2].max())))
            return [np.float32(max(zmax, float(np.float32(c.domposmin[2]))))] + [np.float32(v) for v in g["p0"]]
        # force: pressure force of the fluid on the particles of the block
        nb = int(sim.div.npbok)
        sel = np.nonzero(((sim.code[:nb] & 0x1fff) == g["code"]) & ((sim.code[:nb] & 0xe000) == 0))[0]
        ace = np.zeros((g["count"], 3), np.float32)
        cte = sim.c
        press = lambda rho: np.float32(cte.cteb) * ((rho / np.float32(cte.rhopzero)) ** np.float32(cte.gamma) - np.float32(1))  # noqa: E731
        for p1 in sel:
            idx = self._neigh(sim, sim.pos[p1])
            if idx.size == 0:
                continue
            dr = (sim.pos[p1] - sim.pos[idx]).astype(np.float32)
            rr2 = dr[:, 0] * dr[:, 0] + dr[:, 1] * dr[:, 1] + dr[:, 2] * dr[:, 2]
            ok = (rr2 <= np.float32(cte.kernelsize2)) & (rr2 >= np.float32(1e-18)) & ((sim.code[idx] & 0x1800) == 0x1800)
            idx, dr, rr2 = idx[ok], dr[ok], rr2[ok]
            qq = np.sqrt(rr2) / np.float32(cte.kernelh)
            w1 = np.float32(1) - np.float32(0.5) * qq
            fac = np.float32(cte.bwenh) * w1 * w1 * w1                    # Wendland gradient also in cubic runs
            rho1, rho2 = sim.velrho[p1, 3], sim.velrho[idx, 3]
            pvpm = -((press(rho1) + press(rho2)) / (rho1 * rho2)) * np.float32(cte.massfluid)
            ace[int(sim.idp[p1]) - g["begin"]] = ((pvpm * fac)[:, None] * dr).astype(np.float32).sum(axis=0, dtype=np.float32)
        f = ace.sum(axis=0, dtype=np.float32) * np.float32(cte.massbound)
        return [np.float32(np.sqrt(float(f[0]) ** 2 + float(f[1]) ** 2 + float(f[2]) ** 2)), f[0], f[1], f[2]]

    def update(self, sim, t):
        """JGaugeSystem::CalculeCpu at time t (JDsGaugeSystem.cpp) with JGaugeItem::Update / Output."""
        for g in self.items:
            if not (t >= g["cnext"] and g["cs"] <= t <= g["ce"]):
                continue
            if g["cdt"]:
                nt = int(t / g["cdt"]); g["cnext"] = g["cdt"] * nt
                if g["cnext"] <= t:
                    g["cnext"] = g["cdt"] * (nt + 1)
            res = self.measure(sim, g)
            if g["out"] and t >= g["onext"] and g["os"] <= t <= g["oe"]:
                self.rows[g["name"]].append([t] + [float(v) for v in res])
                if g["odt"]:
                    nt = int(t / g["odt"]); g["onext"] = g["odt"] * nt
                    if g["onext"] <= t:
                        g["onext"] = g["odt"] * (nt + 1)


class MotionOracle:
    """Predefined motion of ONE moving block, restated step by step like the reference does it (state carried from step to step:
    active movement, its start/finish, accumulated phase / velocity) -- JMotion::ProcesTime (JMotion.cpp:579-597),
    JMotionObj::ProcesTime (JMotionObj.cpp:446-640), JMotionMovActive (JMotionObj.cpp:40-55,100-118,262-280), JMotionPos
    (JMotionPos.cpp:45-58), JMatrix4::MatrixRot (JMatrix4.h:465-499).  Single-level objects, movements wait / mvrect / mvrectace /
    mvrot / mvrotace / mvrectsinu / mvrotsinu.  The CUDA path evaluates the same movements in closed form from the time."""

    def __init__(self, objreal):
        self.movs, self.events, self.active = {}, [], []
        for e in objreal:
            if e.tag.startswith("_"):
                continue
            if e.tag == "begin":
                self.events.append([float(e.get("start", 0)), float(e.get("finish")) if e.get("finish") is not None else -1.0, int(e.get("mov"))])
                continue
            deg = e.get("anglesunits", "degrees") == "degrees"
            todeg, torad = (1.0, np.pi / 180.0) if deg else (180.0 / np.pi, 1.0)
            f3 = lambda n: np.array([float(e.find(n).get(a)) for a in "xyz"])  # noqa: E731
            m = {"type": e.tag, "time": float(e.get("duration", 0)), "next": int(e.get("next", 0))}
            if e.tag == "mvrect":
                m["vel"] = f3("vel")
            elif e.tag == "mvrectace":
                m["ace"], m["vel"] = f3("ace"), f3("velini")
            elif e.tag in ("mvrot", "mvrotace", "mvrotsinu"):
                m["p1"], m["p2"] = f3("axisp1"), f3("axisp2")
                if e.tag == "mvrot":
                    m["velang"] = float(e.find("vel").get("ang")) * todeg
                elif e.tag == "mvrotace":
                    m["aceang"], m["velang"] = float(e.find("ace").get("ang")) * todeg, float(e.find("velini").get("ang")) * todeg
                else:
                    m["freq"], m["ampl"] = float(e.find("freq").get("v")), float(e.find("ampl").get("v")) * todeg
                    m["phase"] = float(e.find("phase").get("v")) * torad if e.find("phase") is not None else 0.0
            elif e.tag == "mvrectsinu":
                m["freq"], m["ampl"] = f3("freq"), f3("ampl")
                m["phase"] = f3("phase") * torad if e.find("phase") is not None else np.zeros(3)
            elif e.tag != "wait":
                raise ValueError(f"motion <{e.tag}> is not restated")
            self.movs[int(e.get("id"))] = m
        self.events.sort(key=lambda v: v[0])

    @staticmethod
    def _matrix_rot(ang, p1, p2):
        rad = ang * np.pi / 180.0
        v = p2 - p1
        L, L1 = np.sqrt((v * v).sum()), np.sqrt(v[1] * v[1] + v[2] * v[2])
        T, Tm = np.eye(4), np.eye(4)
        T[:3, 3], Tm[:3, 3] = p1, -p1
        rx, rxm = np.eye(4), np.eye(4)
        if L1 == 0:
            rx[1, 1], rx[1, 2], rx[2, 1], rx[2, 2] = 0, 1, -1, 0
            rxm[1, 1], rxm[1, 2], rxm[2, 1], rxm[2, 2] = 0, -1, 1, 0
        else:
            rx[1, 1], rx[1, 2], rx[2, 1], rx[2, 2] = v[2] / L1, v[1] / L1, -v[1] / L1, v[2] / L1
            rxm[1, 1], rxm[1, 2], rxm[2, 1], rxm[2, 2] = v[2] / L1, -v[1] / L1, v[1] / L1, v[2] / L1
        ry, rym = np.eye(4), np.eye(4)
        ry[0, 0], ry[0, 2], ry[2, 0], ry[2, 2] = L1 / L, v[0] / L, -v[0] / L, L1 / L
        rym[0, 0], rym[0, 2], rym[2, 0], rym[2, 2] = L1 / L, -v[0] / L, v[0] / L, L1 / L
        cs, sn = np.cos(rad), np.sin(rad)
        rz = np.eye(4)
        rz[0, 0], rz[0, 1], rz[1, 0], rz[1, 1] = cs, sn, -sn, cs
        return T @ rx @ ry @ rz @ rym @ rxm @ Tm

    def _activate(self, start, evfinish, movid):
        m = self.movs[movid]
        fin = start + m["time"]
        if evfinish >= 0 and evfinish < fin:
            fin = evfinish
        a = {"id": movid, "start": start, "finish": fin, "evfinish": evfinish, "del": False}
        self._config(a)
        return a

    def _config(self, a):
        m = self.movs[a["id"]]
        a["vel"] = np.array(m.get("vel", np.zeros(3)), np.float64).copy() if m["type"] in ("mvrect", "mvrectace") else np.zeros(3)
        a["velang"] = m.get("velang", 0.0)
        a["phase"] = np.array(m["phase"], np.float64).copy() if m["type"] == "mvrectsinu" else np.zeros(3)
        a["phaseuni"] = m["phase"] if m["type"] == "mvrotsinu" else 0.0

    def proces_time(self, timestep, dt):
        """Returns None (nothing moves), ('lin', d[3]) or ('mat', M[4,4]) for the step [timestep, timestep+dt]."""
        while self.events and self.events[0][0] < timestep + dt:
            st, fin, mv = self.events.pop(0)
            self.active.append(self._activate(st, fin, mv))
        tfin = timestep + dt
        simple, mat, modif = np.zeros(3), None, False
        for a in list(self.active):
            if a["del"]:
                self.active.remove(a)
                continue
            dt2, ts2 = dt, timestep
            while True:
                rep = False
                m = self.movs[a["id"]]
                dtmov = (a["finish"] - ts2) if tfin > a["finish"] else dt2
                dtover = dt2 - dtmov
                if ts2 < a["start"]:
                    dtmov -= (a["start"] - ts2)
                if dtmov > 0:
                    t = dtmov
                    d, ang = None, None
                    if m["type"] == "mvrect":
                        d = m["vel"] * t
                    elif m["type"] == "mvrectace":
                        at = m["ace"] * t
                        d = a["vel"] * t + 0.5 * at * t
                        a["vel"] = a["vel"] + at
                    elif m["type"] == "mvrot":
                        ang = m["velang"] * t
                    elif m["type"] == "mvrotace":
                        at = m["aceang"] * t
                        ang = a["velang"] * t + 0.5 * at * t
                        a["velang"] += at
                    elif m["type"] == "mvrectsinu":
                        ph = a["phase"].copy()
                        p1 = m["ampl"] * np.sin(ph)
                        ph = ph + m["freq"] * (2 * np.pi) * t
                        p2 = m["ampl"] * np.sin(ph)
                        d = np.where(m["ampl"] != 0, p2 - p1, 0.0)
                        a["phase"] = np.where(m["ampl"] != 0, ph, a["phase"])
                    elif m["type"] == "mvrotsinu":
                        ph = a["phaseuni"]
                        a0 = m["ampl"] * np.sin(ph)
                        ph += m["freq"] * (2 * np.pi) * t
                        ang = m["ampl"] * np.sin(ph) - a0
                        a["phaseuni"] = ph
                    if d is not None:                       # JMotionPos::Move
                        if mat is None:
                            simple = simple + d
                        else:
                            mv = np.eye(4); mv[:3, 3] = d
                            mat = mat @ mv
                        modif = True
                    if ang is not None:                     # JMotionPos::Rotate
                        if mat is None:
                            mat = np.eye(4); mat[:3, 3] = simple
                        mat = mat @ self._matrix_rot(ang, m["p1"], m["p2"])
                        modif = True
                if (dtover > 0 or dtmov == 0) and m["next"] in self.movs:
                    a["start"] += m["time"]                 # JMotionMovActive::NextMov
                    a["id"] = m["next"]
                    a["finish"] = a["start"] + self.movs[a["id"]]["time"]
                    if a["evfinish"] >= 0 and a["evfinish"] < a["finish"]:
                        a["finish"] = a["evfinish"]
                    self._config(a)
                    dt2, ts2 = dtover, a["start"]
                    if ts2 <= a["finish"]:
                        rep = True
                if not rep:
                    break
            if tfin > a["finish"]:
                a["del"] = True
        if not modif:
            return None
        return ("lin", simple) if mat is None else ("mat", mat)


def damping_zones(xml_damping) -> np.ndarray:
    """<special><damping> (an ElementTree element) -> (nzones, 64) doubles for orc_damping: the derived values of
    JDsDampingOp_Plane::ReadXml / ComputeDomPlanes and JDsDampingOp_Box::ReadXml (JDsDamping.cpp:58-131,345-391)."""
    out = []
    if xml_damping is None:
        return np.zeros((0, 64))
    f3 = lambda e: np.array([float(e.get("x")), float(e.get("y")), float(e.get("z"))])  # noqa: E731
    for e in xml_damping:
        if e.tag.startswith("_") or e.get("active", "true").lower() in ("false", "0"):
            continue
        z = np.zeros(64)
        z[1] = np.float32(e.find("overlimit").get("value"))
        z[2] = np.float32(e.find("redumax").get("value")) if e.find("redumax") is not None else 10.0
        fx = e.find("factorxyz")
        z[3:6] = [np.float32(fx.get(a, 1)) for a in "xyz"] if fx is not None else 1.0
        if e.tag == "dampingzone":
            lmin, lmax = f3(e.find("limitmin")), f3(e.find("limitmax"))
            vec = lmax - lmin
            dist = np.sqrt((vec * vec).sum())
            u = vec / dist
            z[6:9], z[9], z[10] = u, -(u * lmin).sum(), np.float32(dist)
            dom = e.find("domain")
            if dom is not None:
                z[11], z[12], z[13] = 1, float(dom.get("zmin")), float(dom.get("zmax"))
                vp = [np.array([float(dom.find(f"point{k}").get("x")), float(dom.find(f"point{k}").get("y"))]) for k in (1, 2, 3, 4)]
                ang = [0.0] + [np.degrees(np.arctan2(vp[0][1] - vp[c][1], vp[0][0] - vp[c][0])) % 360.0 for c in (1, 2, 3)]
                order = [0] + sorted((1, 2, 3), key=lambda c: ang[c])
                vp = [vp[c] for c in order]
                for c in range(4):
                    a, b = vp[c], vp[(c + 1) % 4]
                    p1, p2, p3 = np.array([a[0], a[1], 0.0]), np.array([b[0], b[1], 0.0]), np.array([b[0], b[1], 1.0])
                    v = np.cross(p2 - p1, p3 - p1)
                    z[14 + 4 * c:18 + 4 * c] = [v[0], v[1], v[2], -(v * p1).sum()]
        elif e.tag == "dampingbox":
            m1, m2 = f3(e.find("limitmin").find("pointini")), f3(e.find("limitmin").find("pointend"))
            x1, x2 = f3(e.find("limitmax").find("pointini")), f3(e.find("limitmax").find("pointend"))
            dirs = e.find("directions").get("value", "").replace(",", " ").lower().split() if e.find("directions") is not None else []
            bits = {"top": 1, "bottom": 2, "left": 4, "right": 8, "front": 16, "back": 32, "all": 63}
            d = 63
            if dirs:
                d = 0
                for t in dirs:
                    add = not t.startswith("-")
                    t = t.lstrip("+-")
                    d = (d | bits[t]) if add else (d & (63 ^ bits[t]))
            o1, o2 = x1 - z[1], x2 + z[1]
            low, high = (4, 16, 2), (8, 32, 1)      # x: left/right, y: front/back, z: bottom/top
            for k in range(3):
                if not d & high[k]:
                    x2[k] = o2[k] = m2[k]
                if not d & low[k]:
                    x1[k] = o1[k] = m1[k]
            b1, b2 = m1 - x1, x2 - m2
            for k in range(3):
                if not d & high[k] or b2[k] < 0:
                    b2[k] = 0
                if not d & low[k] or b1[k] < 0:
                    b1[k] = 0
            z[0] = 1
            z[6:9], z[9:12], z[12:15], z[15:18], z[18:21], z[21:24], z[24:27], z[27:30] = m1, m2, x1, x2, o1, o2, b1, b2
        else:
            raise ValueError(f"damping <{e.tag}> is not restated")
        out.append(z)
    return np.array(out).reshape(-1, 64)


def run_ref_probe(casepath: str, dirout: str, nsteps: int, options=(), dumplevel: int = 2, threads: int | None = None) -> str:
    """Runs the compiled reference through oracle/_ref/dsph_ref_probe; returns the probe dump directory."""
    if not os.path.exists(REF_PROBE):
        raise FileNotFoundError(REF_PROBE)
    cmd = [REF_PROBE, casepath, dirout, str(nsteps), f"-dumplevel:{dumplevel}", "-sv:none"] + list(options)
    if threads:
        cmd.append(f"-ompthreads:{threads}")
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if out.returncode != 0 or "PROBE_RESULT" not in out.stdout:
        raise RuntimeError("reference probe failed:\n" + out.stdout[-3000:])
    return os.path.join(dirout, "probe")
