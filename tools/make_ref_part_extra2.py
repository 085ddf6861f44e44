#!/usr/bin/env python
"""Generates tests/golden/ref_part_extra2/: PartFloatInfo2.ibi4 and PartMotionRef2.ibi4 as the UNMODIFIED reference CPU build writes
them with -toutx:0.002 (the extra-frequency output of the body data: JSph.cpp:789-790,859-870, JDsPartFloatSave.cpp:189-199,
JDsPartMotionSave.cpp:226-245, JPartMotRefBi4Save.cpp:225-278) for the small dam break with a piston and a floating box of
tests/golden/ref_part_motion.  Needs oracle/_ref (make -C oracle ref)."""
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dualsphysics_b200 import casegen  # noqa: E402

SPEC = dict(dp=0.05, floating_box=(0.15, 0.25, 0.2, 0.1, 0.15, 0.1), floating_rel_density=0.5, piston=True,
            params={"TimeMax": 0.02, "TimeOut": 0.005})
TOUTX = 0.002

if __name__ == "__main__":
    dst = os.path.join(ROOT, "tests", "golden", "ref_part_extra2")
    os.makedirs(dst, exist_ok=True)
    with tempfile.TemporaryDirectory() as td:
        casegen.write_case(td, "Case", casegen.DamBreakSpec(**SPEC))
        exe = os.path.join(ROOT, "oracle", "_ref", "DualSPHysics5.4CPU_linux64")
        subprocess.run([exe, os.path.join(td, "Case"), os.path.join(td, "out"), "-cpu", "-sv:binx", "-toutx:%g" % TOUTX], check=True,
                       stdout=subprocess.DEVNULL, env=dict(os.environ, OMP_NUM_THREADS="2"))
        for f in ("PartFloatInfo2.ibi4", "PartMotionRef2.ibi4"):
            shutil.copy(os.path.join(td, "out", "data", f), os.path.join(dst, f))
    json.dump({"spec": SPEC, "toutx": TOUTX, "reference": "DualSPHysics v5.4 CPU (oracle/_ref)"},
              open(os.path.join(dst, "fixture_extra2.json"), "w"), indent=1)
