"""ctypes binding of libb200sph (include/b2s.h, include/b2s_host.h).

Host-side mirror of the reference's kernel-namespace interface for the WCSPH step: the names and
argument meaning follow JSphGpuSingle / cusph / cudiv (see the citations in include/b2s.h).  The
library is loaded from ``dualsphysics_b200/lib/libb200sph.so`` (built in-tree by
``__graft_entry__.build()``); there is no Python or CPU fallback: if the library or a CUDA device
is missing every call fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2S_LIB_PATH: developer override used by the kernel-tuning scripts (tools/gpu_job.sh) to load an alternative build
LIB_PATH = os.environ.get("B2S_LIB_PATH") or os.path.join(_HERE, "lib", "libb200sph.so")

# enums (same numeric values as the reference)
KERNEL_CUBIC, KERNEL_WENDLAND = 1, 2
STEP_VERLET, STEP_SYMPLECTIC = 1, 2
VISCO_ARTIFICIAL, VISCO_LAMINARSPS, VISCO_LAMINAR = 1, 2, 3
DDT_NONE, DDT_MOLTENI, DDT_FOURTAKAS, DDT_FOURTAKASFULL = 0, 1, 2, 3
BC_DBC, BC_MDBC = 1, 2
INTER_VERLET, INTER_SYMPREDICTOR, INTER_SYMCORRECTOR = 1, 2, 3

ERR_NODEVICE = -2
ERR_UNSUPPORTED = -5


class Params(C.Structure):
    _fields_ = [
        ("simulate2d", C.c_int32), ("tkernel", C.c_int32), ("tstep", C.c_int32), ("tvisco", C.c_int32),
        ("tdensity", C.c_int32), ("tboundary", C.c_int32), ("slipmode", C.c_int32),
        ("verletsteps", C.c_int32), ("cellmode_half", C.c_int32), ("dt_allparticles", C.c_int32),
        ("mdbc_corrector", C.c_int32), ("refsem", C.c_int32), ("use_floating", C.c_int32),
        ("shiftmode", C.c_int32), ("shiftcoef", C.c_float), ("shifttfs", C.c_float),
        ("kernelh", C.c_float), ("cteb", C.c_float), ("gamma", C.c_float), ("rhopzero", C.c_float),
        ("massfluid", C.c_float), ("massbound", C.c_float),
        ("visco", C.c_float), ("viscoboundfactor", C.c_float), ("ddtvalue", C.c_float),
        ("rhopoutmin", C.c_float), ("rhopoutmax", C.c_float), ("gravity", C.c_float * 3),
        ("dp", C.c_double), ("cflnumber", C.c_double), ("dtini", C.c_double), ("dtmin", C.c_double),
        ("coefdtmin", C.c_double), ("ftpause", C.c_double),
        ("maprealposmin", C.c_double * 3), ("maprealposmax", C.c_double * 3),
        ("casenp", C.c_uint32), ("casenfixed", C.c_uint32), ("casenmoving", C.c_uint32),
        ("casenfloat", C.c_uint32), ("casenfluid", C.c_uint32),
        ("periactive", C.c_int32), ("perixinc", C.c_double * 3), ("periyinc", C.c_double * 3), ("perizinc", C.c_double * 3),
        # derived
        ("kernelsize", C.c_float), ("kernelsize2", C.c_float), ("awen", C.c_float), ("bwenh", C.c_float),
        ("cubic_a1", C.c_float), ("cubic_a2", C.c_float), ("cubic_aa", C.c_float), ("cubic_a24", C.c_float),
        ("cubic_c1", C.c_float), ("cubic_d1", C.c_float), ("cubic_c2", C.c_float), ("cubic_od_wdeltap", C.c_float),
        ("eta2", C.c_float), ("ddtkh", C.c_float), ("ddtgz", C.c_float), ("spssmag", C.c_float),
        ("spsblin", C.c_float), ("scell", C.c_float), ("movlimit", C.c_float), ("ovrhopzero", C.c_float),
        ("cs0", C.c_double),
        ("map_posmin", C.c_double * 3), ("map_posmax", C.c_double * 3), ("map_size", C.c_double * 3),
        ("maprealsize", C.c_double * 3),
        ("map_cells", C.c_uint32 * 3), ("domcells", C.c_uint32 * 3), ("domcellcode", C.c_uint32),
        ("scelldiv", C.c_int32),
        ("domposmin", C.c_double * 3), ("domrealposmin", C.c_double * 3), ("domrealposmax", C.c_double * 3),
        ("poscellsize", C.c_float), ("casenbound", C.c_uint32), ("casenpb", C.c_uint32),
    ]


class Floating(C.Structure):
    _fields_ = [
        ("begin", C.c_uint32), ("count", C.c_uint32), ("mass", C.c_float), ("massp", C.c_float),
        ("radius", C.c_float), ("constraints", C.c_int32), ("center", C.c_double * 3),
        ("angles", C.c_float * 3), ("fvel", C.c_float * 3), ("fomega", C.c_float * 3),
        ("facelin", C.c_float * 3), ("faceang", C.c_float * 3), ("inertiaini", C.c_float * 9),
        ("fluforcelin", C.c_float * 3), ("fluforceang", C.c_float * 3), ("preacelin", C.c_float * 3), ("preaceang", C.c_float * 3),
    ]


class MotionRec(C.Structure):
    _fields_ = [
        ("idbegin", C.c_uint32), ("count", C.c_uint32), ("type", C.c_int32), ("pad", C.c_int32),
        ("tstart", C.c_double), ("tend", C.c_double),
        ("v1", C.c_double * 3), ("v2", C.c_double * 3), ("v3", C.c_double * 3),
        ("axisp1", C.c_double * 3), ("axisp2", C.c_double * 3),
    ]


class Info(C.Structure):
    _fields_ = [
        ("np", C.c_uint32), ("npb", C.c_uint32), ("npbok", C.c_uint32), ("npf_out_total", C.c_uint32),
        ("ncx", C.c_uint32), ("ncy", C.c_uint32), ("ncz", C.c_uint32), ("nct", C.c_uint32),
        ("cellmin", C.c_uint32 * 3), ("cellmax", C.c_uint32 * 3),
        ("npb_ignore", C.c_uint32), ("npf_out_last", C.c_uint32), ("npb_out_last", C.c_uint32), ("nstep", C.c_uint32),
        ("timestep", C.c_double), ("dt_last", C.c_double), ("dt_p", C.c_double), ("dt_c", C.c_double),
        ("symplectic_dtpre", C.c_double), ("velmax", C.c_double), ("acemax", C.c_double),
        ("viscdtmax", C.c_float), ("pad0", C.c_float), ("verletstep", C.c_uint32), ("dtmodif", C.c_uint32),
        ("pairs_checked", C.c_uint64), ("pairs_real", C.c_uint64),
    ]


# Every symbol include/b2s.h + include/b2s_host.h declare (checked by tests/test_abi.py).
ABI_SYMBOLS = [
    "b2s_config_finalize", "b2s_error_string", "b2s_create", "b2s_destroy", "b2s_last_error", "b2s_get_params",
    "b2s_upload_particles", "b2s_download_particles", "b2s_download_particles_byid", "b2s_get_array",
    "b2s_get_info", "b2s_set_time", "b2s_celldiv", "b2s_interaction", "b2s_compute_verlet", "b2s_compute_symplectic_pre",
    "b2s_compute_symplectic_corr", "b2s_compute_step", "b2s_run_steps", "b2s_add_motion_rectsinu", "b2s_add_motion", "b2s_add_damping_plane", "b2s_add_damping_box", "b2s_out_enable", "b2s_out_fetch", "b2s_gauge_add_force", "b2s_gauge_add_maxz",
    "b2s_move_linear", "b2s_set_floatings", "b2s_get_floatings", "b2s_get_pos_byid", "b2s_slab_gauges_compute", "b2s_count_pairs", "b2s_time_interaction",
    "b2s_gauge_add_velocity", "b2s_gauge_add_swl", "b2s_gauges_compute", "b2s_gauge_result",
    "b2s_launch_count", "b2s_get_stream", "b2s_set_interaction_variant", "b2s_debug_set_tile_cap", "b2s_slab_config",
    "b2s_slab_histogram", "b2s_slab_set_range", "b2s_slab_drop_ghosts", "b2s_slab_record_bytes", "b2s_slab_pack", "b2s_slab_unpack", "b2s_slab_pack_both", "b2s_slab_unpack_both", "b2s_slab_dt_buffer", "b2s_slab_forces",
    "b2s_slab_apply_dt", "b2s_slab_integrate", "b2s_slab_step_end", "b2s_slab_ft_sums", "b2s_slab_ft_apply", "b2s_slab_get_range",
    "b2s_slab_comm_id", "b2s_slab_comm_init", "b2s_slab_comm_init_local", "b2s_slab_comm_free", "b2s_slab_comm_set_recut", "b2s_slab_comm_stats", "b2s_slab_comm_profile",
    "b2s_slab_exchange", "b2s_slab_step", "b2s_slab_run_steps", "b2s_slab_recut",
    "b2s_sim_load", "b2s_sim_load_slab", "b2s_sim_slab_connect", "b2s_sim_slab_info", "b2s_sim_free", "b2s_sim_ctx", "b2s_sim_params", "b2s_sim_stats", "b2s_sim_set_dirout", "b2s_sim_save_part", "b2s_sim_run",
    "b2s_host_motionref_write", "b2s_host_xml_stats", "b2s_host_last_error",
]

_lib = None


class B2SError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libb200sph error {code}: {msg}")
        self.code = code


def load_library():
    """Loads libb200sph.so (raises if it has not been built: no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} not found: run __graft_entry__.build() (the CUDA library has no fallback)")
    lib = C.CDLL(LIB_PATH)
    vp, u32p = C.c_void_p, C.POINTER(C.c_uint32)
    lib.b2s_error_string.restype = C.c_char_p
    lib.b2s_last_error.restype = C.c_char_p
    lib.b2s_last_error.argtypes = [vp]
    lib.b2s_host_last_error.restype = C.c_char_p
    lib.b2s_config_finalize.argtypes = [C.POINTER(Params)]
    lib.b2s_create.argtypes = [C.POINTER(Params), C.c_int, C.c_uint32, C.POINTER(vp)]
    lib.b2s_destroy.argtypes = [vp]
    lib.b2s_get_params.argtypes = [vp, C.POINTER(Params)]
    lib.b2s_upload_particles.argtypes = [vp, C.c_uint32, vp, vp, vp, vp, vp]
    lib.b2s_download_particles.argtypes = [vp, u32p, vp, vp, vp, vp]
    lib.b2s_download_particles_byid.argtypes = [vp, u32p, vp, vp, vp, vp]
    lib.b2s_get_array.argtypes = [vp, C.c_char_p, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.b2s_get_info.argtypes = [vp, C.POINTER(Info)]
    lib.b2s_set_time.argtypes = [vp, C.c_double, C.c_double]
    lib.b2s_celldiv.argtypes = [vp]
    lib.b2s_interaction.argtypes = [vp, C.c_int]
    lib.b2s_compute_verlet.argtypes = [vp, C.c_double]
    lib.b2s_compute_symplectic_pre.argtypes = [vp, C.c_double]
    lib.b2s_compute_symplectic_corr.argtypes = [vp, C.c_double]
    lib.b2s_compute_step.argtypes = [vp, C.POINTER(C.c_double)]
    lib.b2s_run_steps.argtypes = [vp, C.c_uint32, C.POINTER(C.c_double)]
    lib.b2s_add_motion_rectsinu.argtypes = [vp, C.c_uint32, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                            C.POINTER(C.c_double), C.c_double, C.c_double]
    dp_, fp_ = C.POINTER(C.c_double), C.POINTER(C.c_float)
    lib.b2s_add_damping_plane.argtypes = [vp, dp_, dp_, C.c_float, C.c_float, fp_, dp_]
    lib.b2s_add_damping_box.argtypes = [vp, dp_, dp_, dp_, dp_, C.c_float, C.c_float, fp_, C.c_uint32]
    lib.b2s_out_enable.argtypes = [vp, C.c_uint32]
    lib.b2s_out_fetch.argtypes = [vp, C.c_uint32, u32p, vp, vp, vp, vp, vp]
    lib.b2s_add_motion.argtypes = [vp, C.POINTER(MotionRec)]
    lib.b2s_move_linear.argtypes = [vp, C.c_uint32, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_float)]
    lib.b2s_set_floatings.argtypes = [vp, C.c_uint32, C.POINTER(Floating)]
    lib.b2s_get_floatings.argtypes = [vp, C.c_uint32, C.POINTER(Floating)]
    lib.b2s_get_pos_byid.argtypes = [vp, C.c_uint32, vp, vp]
    lib.b2s_count_pairs.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.b2s_time_interaction.argtypes = [vp, C.c_int, C.POINTER(C.c_float)]
    lib.b2s_launch_count.argtypes = [vp, C.POINTER(C.c_uint64), C.c_int]
    lib.b2s_get_stream.argtypes = [vp, C.POINTER(vp)]
    lib.b2s_set_interaction_variant.argtypes = [vp, C.c_int]
    lib.b2s_debug_set_tile_cap.argtypes = [vp, C.c_uint32]
    lib.b2s_gauge_add_velocity.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    lib.b2s_gauge_add_swl.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.c_float, C.POINTER(C.c_int)]
    lib.b2s_gauge_add_force.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_int)]
    lib.b2s_gauge_add_maxz.argtypes = [vp, C.POINTER(C.c_double), C.c_double, C.c_float, C.POINTER(C.c_int)]
    lib.b2s_gauges_compute.argtypes = [vp]
    lib.b2s_slab_gauges_compute.argtypes = [vp]
    lib.b2s_gauge_result.argtypes = [vp, C.c_int, C.POINTER(C.c_float)]
    lib.b2s_slab_config.argtypes = [vp, C.c_int, C.c_int, C.c_uint32, C.c_uint32]
    lib.b2s_slab_histogram.argtypes = [vp, vp]
    lib.b2s_slab_set_range.argtypes = [vp, C.c_uint32, C.c_uint32]
    lib.b2s_slab_drop_ghosts.argtypes = [vp]
    lib.b2s_slab_record_bytes.argtypes = [vp, C.c_int, u32p]
    lib.b2s_slab_pack.argtypes = [vp, C.c_int, vp, vp, C.c_uint32, u32p, u32p]
    lib.b2s_slab_unpack.argtypes = [vp, C.c_int, vp, C.c_uint32]
    lib.b2s_slab_pack_both.argtypes = [vp, vp, vp, C.c_uint64, u32p]
    lib.b2s_slab_unpack_both.argtypes = [vp, vp, C.c_uint32, C.c_uint32]
    lib.b2s_slab_dt_buffer.argtypes = [vp, C.POINTER(vp)]
    lib.b2s_slab_forces.argtypes = [vp, C.c_int]
    lib.b2s_slab_apply_dt.argtypes = [vp, C.c_int]
    lib.b2s_slab_integrate.argtypes = [vp, C.c_int]
    lib.b2s_slab_step_end.argtypes = [vp, C.POINTER(C.c_double)]
    lib.b2s_slab_ft_sums.argtypes = [vp, C.POINTER(vp), u32p]
    lib.b2s_slab_ft_apply.argtypes = [vp, C.c_int]
    lib.b2s_sim_load_slab.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    lib.b2s_sim_slab_info.argtypes = [vp, u32p, C.c_uint32, u32p]
    lib.b2s_sim_load.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_uint32, C.POINTER(vp)]
    lib.b2s_sim_free.argtypes = [vp]
    lib.b2s_sim_ctx.argtypes = [vp, C.POINTER(vp)]
    lib.b2s_slab_get_range.argtypes = [vp, u32p]
    lib.b2s_slab_comm_id.argtypes = [C.c_void_p]
    lib.b2s_slab_comm_init.argtypes = [vp, C.c_void_p, C.c_uint32]
    lib.b2s_slab_comm_init_local.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_uint32]
    lib.b2s_slab_comm_free.argtypes = [vp]
    lib.b2s_slab_comm_set_recut.argtypes = [vp, C.c_uint32]
    lib.b2s_slab_comm_stats.argtypes = [vp, C.POINTER(C.c_double), C.c_int]
    lib.b2s_slab_comm_profile.argtypes = [vp, C.c_char_p, C.c_uint32]
    lib.b2s_slab_exchange.argtypes = [vp]
    lib.b2s_slab_step.argtypes = [vp, C.POINTER(C.c_double)]
    lib.b2s_slab_run_steps.argtypes = [vp, C.c_uint32, C.POINTER(C.c_double)]
    lib.b2s_slab_recut.argtypes = [vp, u32p, C.c_uint32]
    lib.b2s_sim_slab_connect.argtypes = [vp, C.c_void_p]
    lib.b2s_sim_params.argtypes = [vp, C.POINTER(Params)]
    lib.b2s_sim_save_part.argtypes = [vp]
    lib.b2s_sim_set_dirout.argtypes = [vp, C.c_char_p]
    lib.b2s_host_xml_stats.argtypes = [C.c_char_p, u32p, u32p]
    lib.b2s_sim_stats.argtypes = [vp, u32p, u32p]
    lib.b2s_sim_run.argtypes = [vp, C.c_uint32, C.POINTER(C.c_double), u32p]
    lib.b2s_host_motionref_write.argtypes = [C.c_char_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                             C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    _lib = lib
    return lib


_ARRAY_SPEC = {  # name -> (dtype, components)
    "idp": (np.uint32, 1), "code": (np.uint16, 1), "dcell": (np.uint32, 1), "pos": (np.float64, 3),
    "velrho": (np.float32, 4), "velrhopre": (np.float32, 4), "velrhom1": (np.float32, 4), "pospre": (np.float64, 3),
    "cellpart": (np.uint32, 1), "sortpart": (np.uint32, 1), "begincell": (np.uint32, 1), "beginendcell": (np.int32, 2),
    "ace": (np.float32, 3), "ar": (np.float32, 1), "press": (np.float32, 1), "poscell": (np.float32, 4),
    "posp": (np.float32, 4), "acear": (np.float32, 4), "spstau": (np.float32, 6), "boundnor": (np.float32, 4),
    "ridp": (np.uint32, 1),
}


class Context:
    """Thin object wrapper over a ``b2s_ctx*`` (one per GPU)."""

    def __init__(self, handle, lib, owner=None):
        self._h, self._lib, self._owner = handle, lib, owner

    @classmethod
    def create(cls, params: Params, device: int = -1, capacity: int = 0):
        lib = load_library()
        h = C.c_void_p()
        r = lib.b2s_create(C.byref(params), device, capacity, C.byref(h))
        if r:
            raise B2SError(r, lib.b2s_error_string(r).decode())
        return cls(h, lib)

    def _chk(self, r):
        if r:
            raise B2SError(r, self._lib.b2s_last_error(self._h).decode() or self._lib.b2s_error_string(r).decode())

    def close(self):
        if self._h and self._owner is None:
            self._lib.b2s_destroy(self._h)
        self._h = None

    @property
    def params(self) -> Params:
        p = Params()
        self._chk(self._lib.b2s_get_params(self._h, C.byref(p)))
        return p

    def upload(self, idp, code, pos, velrho, boundnor=None):
        idp = np.ascontiguousarray(idp, np.uint32)
        code = np.ascontiguousarray(code, np.uint16)
        pos = np.ascontiguousarray(pos, np.float64)
        velrho = np.ascontiguousarray(velrho, np.float32)
        bn = None if boundnor is None else np.ascontiguousarray(boundnor, np.float32)
        self._chk(self._lib.b2s_upload_particles(self._h, idp.shape[0], idp.ctypes.data, code.ctypes.data,
                                                 pos.ctypes.data, velrho.ctypes.data,
                                                 None if bn is None else bn.ctypes.data))

    def upload_ptrs(self, n, idp_ptr, code_ptr, pos_ptr, velrho_ptr, boundnor_ptr=None):
        """Upload from raw host pointers (e.g. pinned torch tensors' data_ptr())."""
        self._chk(self._lib.b2s_upload_particles(self._h, n, idp_ptr, code_ptr, pos_ptr, velrho_ptr, boundnor_ptr))

    def info(self) -> Info:
        i = Info()
        self._chk(self._lib.b2s_get_info(self._h, C.byref(i)))
        return i

    def array(self, name: str) -> np.ndarray:
        dt, k = _ARRAY_SPEC[name]
        inf = self.info()
        if name in ("begincell",):
            n = inf.nct * 2 + 6
        elif name == "beginendcell":
            n = inf.nct * 2 + 5
        elif name == "ridp":
            p = self.params
            n = p.casenmoving + p.casenfloat
        else:
            # cellpart/sortpart cover the particles that ENTERED the divide: + excluded ones and the previous periodic duplicates / ghosts
            n = max(2 * inf.np + inf.npf_out_last + inf.npb_out_last + 16, 1)
        buf = np.empty((n, k) if k > 1 else (n,), dtype=dt)
        cnt = C.c_size_t(0)
        self._chk(self._lib.b2s_get_array(self._h, name.encode(), buf.ctypes.data, buf.nbytes, C.byref(cnt)))
        return buf[:cnt.value]

    def download_byid(self):
        n = self.params.casenp
        idp = np.empty(n, np.uint32); pos = np.empty((n, 3), np.float64)
        vel = np.empty((n, 3), np.float32); rho = np.empty(n, np.float32)
        cnt = C.c_uint32(0)
        self._chk(self._lib.b2s_download_particles_byid(self._h, C.byref(cnt), idp.ctypes.data, pos.ctypes.data,
                                                        vel.ctypes.data, rho.ctypes.data))
        c = cnt.value
        return idp[:c], pos[:c], vel[:c], rho[:c]

    def celldiv(self):
        self._chk(self._lib.b2s_celldiv(self._h))

    def interaction(self, interstep: int):
        self._chk(self._lib.b2s_interaction(self._h, interstep))

    def compute_verlet(self, dt: float):
        self._chk(self._lib.b2s_compute_verlet(self._h, dt))

    def compute_symplectic_pre(self, dt: float):
        self._chk(self._lib.b2s_compute_symplectic_pre(self._h, dt))

    def compute_symplectic_corr(self, dt: float):
        self._chk(self._lib.b2s_compute_symplectic_corr(self._h, dt))

    def compute_step(self) -> float:
        dt = C.c_double(0)
        self._chk(self._lib.b2s_compute_step(self._h, C.byref(dt)))
        return dt.value

    def run_steps(self, n: int) -> float:
        t = C.c_double(0)
        self._chk(self._lib.b2s_run_steps(self._h, n, C.byref(t)))
        return t.value

    def count_pairs(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        self._chk(self._lib.b2s_count_pairs(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def time_interaction(self, reps: int) -> float:
        ms = C.c_float(0)
        self._chk(self._lib.b2s_time_interaction(self._h, reps, C.byref(ms)))
        return ms.value

    def launch_count(self, reset=False) -> int:
        n = C.c_uint64(0)
        self._chk(self._lib.b2s_launch_count(self._h, C.byref(n), 1 if reset else 0))
        return n.value

    def set_variant(self, v: int):
        self._chk(self._lib.b2s_set_interaction_variant(self._h, v))

    def set_tile_cap(self, cap: int):
        self._chk(self._lib.b2s_debug_set_tile_cap(self._h, cap))

    # ---- slab decomposition (include/b2s.h, "slab" section) ----
    def slab_config(self, rank, nranks, cx_begin, cx_end):
        self._chk(self._lib.b2s_slab_config(self._h, rank, nranks, cx_begin, cx_end))

    def slab_histogram(self, dev_ptr):
        self._chk(self._lib.b2s_slab_histogram(self._h, dev_ptr))

    def slab_set_range(self, cx_begin, cx_end):
        self._chk(self._lib.b2s_slab_set_range(self._h, cx_begin, cx_end))

    def slab_get_range(self):
        rng = (C.c_uint32 * 2)()
        self._chk(self._lib.b2s_slab_get_range(self._h, rng))
        return int(rng[0]), int(rng[1])

    def slab_drop_ghosts(self):
        self._chk(self._lib.b2s_slab_drop_ghosts(self._h))

    def slab_record_bytes(self, kind):
        b = C.c_uint32(0)
        self._chk(self._lib.b2s_slab_record_bytes(self._h, kind, C.byref(b)))
        return b.value

    def slab_pack(self, kind, left_ptr, right_ptr, cap_records):
        a, b = C.c_uint32(0), C.c_uint32(0)
        self._chk(self._lib.b2s_slab_pack(self._h, kind, left_ptr, right_ptr, cap_records, C.byref(a), C.byref(b)))
        return a.value, b.value

    def slab_pack_both(self, left_ptr, right_ptr, cap_bytes):
        c = (C.c_uint32 * 4)()
        self._chk(self._lib.b2s_slab_pack_both(self._h, left_ptr, right_ptr, cap_bytes, c))
        return list(c)

    def slab_unpack_both(self, ptr, nmigrants, nhalo):
        self._chk(self._lib.b2s_slab_unpack_both(self._h, ptr, nmigrants, nhalo))

    def slab_unpack(self, kind, ptr, count):
        self._chk(self._lib.b2s_slab_unpack(self._h, kind, ptr, count))

    def slab_dt_buffer(self):
        p = C.c_void_p()
        self._chk(self._lib.b2s_slab_dt_buffer(self._h, C.byref(p)))
        return p.value

    def slab_forces(self, interstep):
        self._chk(self._lib.b2s_slab_forces(self._h, interstep))

    def slab_apply_dt(self, interstep):
        self._chk(self._lib.b2s_slab_apply_dt(self._h, interstep))

    def slab_integrate(self, mode):
        self._chk(self._lib.b2s_slab_integrate(self._h, mode))

    def slab_ft_sums(self):
        p, n = C.c_void_p(), C.c_uint32(0)
        self._chk(self._lib.b2s_slab_ft_sums(self._h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def slab_ft_apply(self, predictor):
        self._chk(self._lib.b2s_slab_ft_apply(self._h, int(predictor)))

    def slab_step_end(self):
        dt = C.c_double(0)
        self._chk(self._lib.b2s_slab_step_end(self._h, C.byref(dt)))
        return dt.value

    # ---- the slab step inside the library (csrc/slab_comm.cu): NCCL transport, no host synchronisation inside a step ----
    def slab_comm_init(self, nccl_id: bytes, max_plane_particles: int):
        assert len(nccl_id) == 128
        self._chk(self._lib.b2s_slab_comm_init(self._h, (C.c_uint8 * 128).from_buffer_copy(nccl_id), max_plane_particles))

    def slab_comm_set_recut(self, every_steps: int):
        self._chk(self._lib.b2s_slab_comm_set_recut(self._h, every_steps))

    def slab_comm_stats(self, reset=False):
        """(ms inside pack+transfer+unpack, exchanges, bytes sent, re-cuts) since the last reset, migrants sent since init."""
        out = (C.c_double * 5)()
        self._chk(self._lib.b2s_slab_comm_stats(self._h, out, int(reset)))
        return out[0], int(out[1]), int(out[2]), int(out[3]), int(out[4])

    def slab_comm_profile(self) -> str:
        buf = C.create_string_buffer(1024)
        self._chk(self._lib.b2s_slab_comm_profile(self._h, buf, 1024))
        return buf.value.decode()

    def slab_exchange(self):
        self._chk(self._lib.b2s_slab_exchange(self._h))

    def slab_step(self) -> float:
        dt = C.c_double(0)
        self._chk(self._lib.b2s_slab_step(self._h, C.byref(dt)))
        return dt.value

    def slab_run_steps(self, n: int) -> float:
        t = C.c_double(0)
        self._chk(self._lib.b2s_slab_run_steps(self._h, n, C.byref(t)))
        return t.value

    def slab_recut(self):
        rng = (C.c_uint32 * 2)()
        self._chk(self._lib.b2s_slab_recut(self._h, rng, 2))
        return int(rng[0]), int(rng[1])

    def stream(self):
        p = C.c_void_p()
        self._chk(self._lib.b2s_get_stream(self._h, C.byref(p)))
        return p.value

    # ---- gauges
    def gauge_add_velocity(self, point):
        gid = C.c_int(0)
        self._chk(self._lib.b2s_gauge_add_velocity(self._h, (C.c_double * 3)(*point), C.byref(gid)))
        return gid.value

    def gauge_add_swl(self, point0, point2, pointdp, masslimit):
        gid = C.c_int(0)
        self._chk(self._lib.b2s_gauge_add_swl(self._h, (C.c_double * 3)(*point0), (C.c_double * 3)(*point2), pointdp, masslimit, C.byref(gid)))
        return gid.value

    def gauge_add_force(self, idbegin, count, typevalue):
        gid = C.c_int(-1)
        self._chk(self._lib.b2s_gauge_add_force(self._h, idbegin, count, typevalue, C.byref(gid)))
        return gid.value

    def gauge_add_maxz(self, point0, height, distlimit):
        gid = C.c_int(-1)
        self._chk(self._lib.b2s_gauge_add_maxz(self._h, (C.c_double * 3)(*point0), height, distlimit, C.byref(gid)))
        return gid.value

    def gauges_compute(self):
        self._chk(self._lib.b2s_gauges_compute(self._h))

    def slab_gauges_compute(self):
        """Gauges of a slab-decomposed run, collectively (b2s_slab_gauges_compute); b2s_slab_step evaluates them after every step."""
        self._chk(self._lib.b2s_slab_gauges_compute(self._h))

    def gauge_result(self, gid):
        out = (C.c_float * 3)()
        self._chk(self._lib.b2s_gauge_result(self._h, gid, out))
        return [out[0], out[1], out[2]]

    def add_damping_plane(self, limitmin, limitmax, overlimit, redumax=10.0, factorxyz=(1, 1, 1), domain=None):
        dom = (C.c_double * 10)(*domain) if domain is not None else None
        self._chk(self._lib.b2s_add_damping_plane(self._h, (C.c_double * 3)(*limitmin), (C.c_double * 3)(*limitmax),
                                                  overlimit, redumax, (C.c_float * 3)(*factorxyz), dom))

    def add_damping_box(self, limitmin_ini, limitmin_end, limitmax_ini, limitmax_end, overlimit, redumax=10.0,
                        factorxyz=(1, 1, 1), directions=63):
        d3 = C.c_double * 3
        self._chk(self._lib.b2s_add_damping_box(self._h, d3(*limitmin_ini), d3(*limitmin_end), d3(*limitmax_ini), d3(*limitmax_end),
                                                overlimit, redumax, (C.c_float * 3)(*factorxyz), directions))

    def out_enable(self, capacity: int):
        self._chk(self._lib.b2s_out_enable(self._h, capacity))

    def out_fetch(self, maxcount: int = 1 << 20):
        idp = np.empty(maxcount, np.uint32); pos = np.empty((maxcount, 3), np.float64); vel = np.empty((maxcount, 3), np.float32)
        rho = np.empty(maxcount, np.float32); mot = np.empty(maxcount, np.uint8)
        cnt = C.c_uint32(0)
        self._chk(self._lib.b2s_out_fetch(self._h, maxcount, C.byref(cnt), idp.ctypes.data, pos.ctypes.data, vel.ctypes.data,
                                          rho.ctypes.data, mot.ctypes.data))
        c = cnt.value
        return idp[:c], pos[:c], vel[:c], rho[:c], mot[:c]

    def set_floatings(self, fts):
        arr = (Floating * len(fts))(*fts)
        self._chk(self._lib.b2s_set_floatings(self._h, len(fts), arr))

    def pos_byid(self, ids) -> np.ndarray:
        """Positions of moving / floating particles by id (b2s_get_pos_byid); NaN rows for particles that left the simulation."""
        ids = np.ascontiguousarray(ids, np.uint32)
        out = np.empty((ids.shape[0], 3), np.float64)
        self._chk(self._lib.b2s_get_pos_byid(self._h, ids.shape[0], ids.ctypes.data, out.ctypes.data))
        return out

    def get_floatings(self, n):
        arr = (Floating * n)()
        self._chk(self._lib.b2s_get_floatings(self._h, n, arr))
        return list(arr)


def motionref_write(dirdata: str, bodies, timeout: float, mkboundfirst: int, idp, pos, timestep, step) -> None:
    """PartMotionRef.ibi4 from host arrays (b2s_host_motionref_write): bodies = [(floating, mkbound, begin, count)], idp[n] ascending,
    pos[nparts, n, 3], timestep[nparts], step[nparts].  Host code only, no GPU needed."""
    lib = load_library()
    fl = np.ascontiguousarray([b[0] for b in bodies], dtype=np.uint8)
    mk = np.ascontiguousarray([b[1] for b in bodies], dtype=np.uint16)
    bg = np.ascontiguousarray([b[2] for b in bodies], dtype=np.uint32)
    ct = np.ascontiguousarray([b[3] for b in bodies], dtype=np.uint32)
    idp = np.ascontiguousarray(idp, dtype=np.uint32)
    pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, len(idp), 3)
    ts = np.ascontiguousarray(timestep, dtype=np.float64)
    st = np.ascontiguousarray(step, dtype=np.uint32)
    d = dirdata if dirdata.endswith("/") else dirdata + "/"
    r = lib.b2s_host_motionref_write(d.encode(), len(bodies), fl.ctypes.data, mk.ctypes.data, bg.ctypes.data, ct.ctypes.data,
                                     float(timeout), int(mkboundfirst), pos.shape[0], len(idp), idp.ctypes.data, pos.ctypes.data,
                                     ts.ctypes.data, st.ctypes.data)
    if r:
        raise B2SError(r, lib.b2s_host_last_error().decode())


def slab_comm_id() -> bytes:
    """The 128-byte NCCL id of a new communicator (rank 0 makes it, every rank passes it to Context.slab_comm_init)."""
    buf = (C.c_uint8 * 128)()
    r = load_library().b2s_slab_comm_id(buf)
    if r:
        raise B2SError(r, "b2s_slab_comm_id: NCCL not available")
    return bytes(buf)


def slab_comm_init_local(ctxs, max_plane_particles: int):
    """Test transport: all ranks of the run live in this process on one device (csrc/slab_comm.cu)."""
    arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
    r = load_library().b2s_slab_comm_init_local(arr, len(ctxs), max_plane_particles)
    if r:
        raise B2SError(r, "b2s_slab_comm_init_local")


class Simulation:
    """The C++ host driver: Case.xml + Case.bi4 in, Part_XXXX.bi4 out (JSphGpuSingle::Run-style)."""

    def __init__(self, casepath: str, dirout: str | None = None, options: str = "", device: int = -1,
                 extra_capacity: int = 0, slab: tuple | None = None):
        """slab=(rank, nranks): load only this rank's slab (b2s_sim_load_slab)."""
        self._lib = load_library()
        self._h = C.c_void_p()
        if slab is not None:
            r = self._lib.b2s_sim_load_slab(casepath.encode(), options.encode(), device, slab[0], slab[1], C.byref(self._h))
        else:
            r = self._lib.b2s_sim_load(casepath.encode(), (dirout or "").encode(), options.encode(), device,
                                       extra_capacity, C.byref(self._h))
        if r:
            raise B2SError(r, self._lib.b2s_host_last_error().decode())
        ch = C.c_void_p()
        self._lib.b2s_sim_ctx(self._h, C.byref(ch))
        self.ctx = Context(ch, self._lib, owner=self)

    @property
    def params(self) -> Params:
        p = Params()
        self._lib.b2s_sim_params(self._h, C.byref(p))
        return p

    def slab_info(self, nranks: int):
        cuts = (C.c_uint32 * (nranks + 1))()
        mp = C.c_uint32(0)
        r = self._lib.b2s_sim_slab_info(self._h, cuts, nranks + 1, C.byref(mp))
        if r:
            raise B2SError(r, "b2s_sim_slab_info")
        return list(cuts), mp.value

    def run(self, nsteps: int = 0):
        t, n = C.c_double(0), C.c_uint32(0)
        r = self._lib.b2s_sim_run(self._h, nsteps, C.byref(t), C.byref(n))
        if r:
            raise B2SError(r, self._lib.b2s_host_last_error().decode())
        return t.value, n.value

    def set_dirout(self, dirout: str):
        """Output directory of a slab rank (b2s_sim_set_dirout): PART files are then written as pieces."""
        r = self._lib.b2s_sim_set_dirout(self._h, dirout.encode())
        if r:
            raise B2SError(r, self._lib.b2s_host_last_error().decode())

    def save_part(self):
        r = self._lib.b2s_sim_save_part(self._h)
        if r:
            raise B2SError(r, self._lib.b2s_host_last_error().decode())

    def close(self):
        if self._h:
            self._lib.b2s_sim_free(self._h)
            self._h = None
            self.ctx._h = None
