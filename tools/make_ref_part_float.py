#!/usr/bin/env python
"""Generates tests/golden/ref_part_float/: the list files the UNMODIFIED reference CPU build writes next to its PART files
(PartInfo.ibi4, PartOut_000.obi4, PartFloatInfo.ibi4) for a small dam break with a piston and a floating box; RhopOutMin is
raised to 999.5 so that particles are excluded in every PART interval.  Needs oracle/_ref (make -C oracle ref)."""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dualsphysics_b200 import casegen  # noqa: E402

SPEC = dict(dp=0.05, floating_box=(0.15, 0.25, 0.2, 0.1, 0.15, 0.1), floating_rel_density=0.5, piston=True,
            params={"TimeMax": 0.02, "TimeOut": 0.005, "RhopOutMin": 999.5})

if __name__ == "__main__":
    dst = os.path.join(ROOT, "tests", "golden", "ref_part_float")
    os.makedirs(dst, exist_ok=True)
    with tempfile.TemporaryDirectory() as td:
        casegen.write_case(td, "Case", casegen.DamBreakSpec(**SPEC))
        exe = os.path.join(ROOT, "oracle", "_ref", "DualSPHysics5.4CPU_linux64")
        subprocess.run([exe, os.path.join(td, "Case"), os.path.join(td, "out"), "-cpu", "-sv:binx"], check=True,
                       stdout=subprocess.DEVNULL, env=dict(os.environ, OMP_NUM_THREADS="2"))
        for f in ("PartFloatInfo.ibi4", "PartInfo.ibi4", "PartOut_000.obi4"):
            shutil.copy(os.path.join(td, "out", "data", f), os.path.join(dst, f))
