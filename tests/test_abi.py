"""CPU-side checks of the drop-in boundary: the library loads, exports every declared symbol, and the
host-only entry points (configuration math, error strings) behave; no compute call is made."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import helpers as H
import dualsphysics_b200 as dsb
from dualsphysics_b200.lib import ABI_SYMBOLS


def _declared_symbols():
    names = []
    for hdr in ("b2s.h", "b2s_host.h"):
        txt = open(os.path.join(H.ROOT, "include", hdr)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names += re.findall(r"\b(b2s_[a-z0-9_]+)\s*\(", txt)
    return sorted(set(names))


def test_header_symbols_match_binding_list():
    assert _declared_symbols() == sorted(set(ABI_SYMBOLS))


def test_library_exports_every_declared_symbol(built_lib):
    missing = [s for s in _declared_symbols() if not hasattr(built_lib, s)]
    assert not missing, missing


def test_struct_sizes_match_c(built_lib, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include "b2s.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu\\n",sizeof(b2s_params),sizeof(b2s_info),sizeof(b2s_floating));return 0;}\n')
    exe = tmp_path / "sz"
    import subprocess
    subprocess.check_call(["gcc", "-I", os.path.join(H.ROOT, "include"), str(src), "-o", str(exe)])
    a, b, c = map(int, subprocess.check_output([str(exe)]).split())
    assert (a, b, c) == (C.sizeof(dsb.Params), C.sizeof(dsb.Info), C.sizeof(dsb.Floating))


def test_config_finalize_matches_reference_constants(built_lib, tmp_path):
    """Derived constants vs the values the compiled reference reports for the same case (golden init dump)."""
    fx = "sym_wend_art_ddt2"
    import json
    meta = json.load(open(os.path.join(H.GOLDEN, fx, "fixture.json")))
    H.write_case(tmp_path, **meta["spec"])
    p, _ = H.params_from_case(str(tmp_path / "Case"))
    g = H.b2sd.load(os.path.join(H.GOLDEN, fx, "s0000_init.b2sd"))
    assert p.domcellcode == int(g["domcellcode"])
    assert [p.map_cells[k] for k in range(3)] == [int(g["mapcells_x"]), int(g["mapcells_y"]), int(g["mapcells_z"])]
    for name, ref in (("kernelsize", "kernelsize"), ("scell", "scell"), ("eta2", "eta2"), ("ddtkh", "ddtkh"),
                      ("ddtgz", "ddtgz"), ("awen", "awen"), ("bwenh", "bwenh"), ("movlimit", "movlimit")):
        assert np.float32(getattr(p, name)) == np.float32(g[ref]), name
    for name, ref in (("cs0", "cs0"), ("dtini", "dtini"), ("dtmin", "dtmin")):
        assert getattr(p, name) == float(g[ref]), name
    for k, ax in enumerate("xyz"):
        assert p.map_posmin[k] == float(g["mapposmin_" + ax])
        assert p.map_posmax[k] == float(g["mapposmax_" + ax])


def test_no_device_fails_loudly(built_lib):
    if H.have_gpu():
        pytest.skip("a CUDA device is present")
    p = dsb.Params()
    p.tkernel, p.tstep, p.tvisco, p.tboundary = 2, 2, 1, 1
    p.kernelh, p.cteb, p.gamma, p.rhopzero, p.dp, p.cflnumber = 0.03, 1e5, 7, 1000, 0.02, 0.2
    p.massfluid = p.massbound = 0.008
    p.visco = 0.1
    p.gravity[2] = -9.81
    for k in range(3):
        p.maprealposmax[k] = 1.0
    p.casenp = p.casenfluid = 10
    with pytest.raises(dsb.B2SError) as e:
        dsb.Context.create(p, capacity=100)
    assert e.value.code == dsb.lib.ERR_NODEVICE
    with pytest.raises(dsb.B2SError):
        meta_dir = os.path.join(H.ROOT, "tests")
        dsb.Simulation(os.path.join(meta_dir, "does_not_exist"))


def test_error_strings(built_lib):
    assert b"no CUDA device" in built_lib.b2s_error_string(-2)
    assert built_lib.b2s_error_string(0) == b"ok"
