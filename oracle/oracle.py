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
    return c


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleSim:
    """Whole-step CPU oracle over numpy arrays (cell-sorted order, like the solver's state)."""

    def __init__(self, cte: Cte, idp, code, pos, velrho, npb: int, dtini: float, verletsteps: int = 40,
                 ftomassp=None, boundnor=None, mdbc_corrector=False):
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
        self.mdbc_corrector = mdbc_corrector
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

    # RunCellDivide (JSphCpuSingle.cpp:497-586)
    def celldiv(self):
        n = self.np_
        r = self.L.orc_celldiv(C.byref(self.c), n, self.npb, _p(self.dcell), _p(self.code), _p(self.cellpart),
                               _p(self.sortpart), _p(self.begincell), C.byref(self.div))
        assert r == 0
        sp = self.sortpart[:n]
        for name in ("idp", "code", "dcell", "pos", "velrho", "velrhom1", "tau", "pospre", "velrhopre", "boundnor"):
            a = getattr(self, name)
            if a is not None:
                setattr(self, name, np.ascontiguousarray(a[:n][sp]))
        self.npf_out_total += int(self.div.npf_out)
        self.np_, self.npb = int(self.div.np), int(self.div.npb)
        if self.div.npb_out:
            raise RuntimeError("oracle: boundary particles excluded")

    def mdbc(self):
        """MdbcBoundCorrection (JSphCpuSingle.cpp:745-760), SlipMode vel0."""
        if self.boundnor is None:
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
                               _p(self.ace), _p(self.ar), _p(self.press), C.byref(vm), C.byref(am), C.byref(vd))
        self.velmax, self.acemax, self.viscdtmax = vm.value, am.value, vd.value
        return self.L.orc_dtvariable(C.byref(self.c), vm.value, am.value, vd.value)

    def step(self) -> float:
        n = self.np_
        if self.tstep == 2:
            dt = self.symplectic_dtpre
            dt_p = self.interaction()
            self.pospre, self.velrhopre = self.pos[:n].copy(), self.velrho[:n].copy()
            self.pos, self.velrho = self.pospre.copy(), self.velrhopre.copy()
            self.L.orc_symplectic_pre(C.byref(self.c), n, self.npb, C.c_double(dt), _p(self.pospre), _p(self.velrhopre),
                                      _p(self.ar), _p(self.ace), _p(self.pos), _p(self.velrho), _p(self.dcell), _p(self.code))
            self.celldiv()
            n = self.np_
            dt_c = self.interaction(corrector=True)
            self.L.orc_symplectic_corr(C.byref(self.c), n, self.npb, C.c_double(dt), _p(self.pospre), _p(self.velrhopre),
                                       _p(self.ar), _p(self.ace), _p(self.pos), _p(self.velrho), _p(self.dcell), _p(self.code))
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
            self.L.orc_verlet(C.byref(self.c), n, self.npb, C.c_double(dt), 1 if euler else 0, _p(self.velrho), _p(v2),
                              _p(self.ar), _p(self.ace), _p(self.pos), _p(self.dcell), _p(self.code), _p(new))
            self.velrhom1 = self.velrho[:n]
            self.velrho = new
        self.celldiv()
        self.timestep += dt
        self.nstep += 1
        return dt


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
