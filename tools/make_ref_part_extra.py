#!/usr/bin/env python
"""Generates tests/golden/ref_part_extra/: PartExtra_0001.bi4 as the UNMODIFIED reference CPU build writes it with -svextraparts:1
for a small mDBC dam break (JDsExtraData.cpp:85-140: the boundary normals by id, already doubled to reach the ghost node), the file a
restart with mDBC needs (JSph::ConfigBoundNormals, JSph.cpp:1396-1403).  Needs oracle/_ref (make -C oracle ref)."""
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dualsphysics_b200 import casegen  # noqa: E402

SPEC = dict(dp=0.05, layers=4, normals=True, params={"Boundary": 2, "SlipMode": 1, "TimeMax": 0.01, "TimeOut": 0.005})

if __name__ == "__main__":
    dst = os.path.join(ROOT, "tests", "golden", "ref_part_extra")
    os.makedirs(dst, exist_ok=True)
    with tempfile.TemporaryDirectory() as td:
        casegen.write_case(td, "Case", casegen.DamBreakSpec(**SPEC))
        exe = os.path.join(ROOT, "oracle", "_ref", "DualSPHysics5.4CPU_linux64")
        subprocess.run([exe, os.path.join(td, "Case"), os.path.join(td, "out"), "-cpu", "-sv:binx", "-svextraparts:1"], check=True,
                       stdout=subprocess.DEVNULL, env=dict(os.environ, OMP_NUM_THREADS="2"))
        print(sorted(os.listdir(os.path.join(td, "out", "data"))))
        shutil.copy(os.path.join(td, "out", "data", "PartExtra_0001.bi4"), os.path.join(dst, "PartExtra_0001.bi4"))
    json.dump({"spec": SPEC, "options": ["-svextraparts:1"], "reference": "DualSPHysics v5.4 CPU (oracle/_ref)"},
              open(os.path.join(dst, "fixture_extra.json"), "w"), indent=1)
