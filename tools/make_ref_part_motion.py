#!/usr/bin/env python
"""Generates tests/golden/ref_part_motion/: PartMotionRef.ibi4 as the UNMODIFIED reference CPU build writes it (three reference
particles per moving / floating body and their positions at every PART; JDsPartMotionSave.cpp, JPartMotRefBi4Save.cpp) for a
small dam break with a piston and a floating box, plus positions and ids of the boundary particles of every PART (from the
reference's Part_XXXX.bi4 written with -saveposdouble:1) as bound.npz, the input of the writer under test.
Needs oracle/_ref (make -C oracle ref)."""
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dualsphysics_b200 import bi4, casegen  # noqa: E402

SPEC = dict(dp=0.05, floating_box=(0.15, 0.25, 0.2, 0.1, 0.15, 0.1), floating_rel_density=0.5, piston=True,
            params={"TimeMax": 0.02, "TimeOut": 0.005})

if __name__ == "__main__":
    dst = os.path.join(ROOT, "tests", "golden", "ref_part_motion")
    os.makedirs(dst, exist_ok=True)
    with tempfile.TemporaryDirectory() as td:
        casegen.write_case(td, "Case", casegen.DamBreakSpec(**SPEC))
        exe = os.path.join(ROOT, "oracle", "_ref", "DualSPHysics5.4CPU_linux64")
        subprocess.run([exe, os.path.join(td, "Case"), os.path.join(td, "out"), "-cpu", "-sv:binx", "-saveposdouble:1"], check=True,
                       stdout=subprocess.DEVNULL, env=dict(os.environ, OMP_NUM_THREADS="2"))
        data = os.path.join(td, "out", "data")
        shutil.copy(os.path.join(data, "PartMotionRef.ibi4"), os.path.join(dst, "PartMotionRef.ibi4"))
        out = {}
        k = 0
        while os.path.exists(os.path.join(data, "Part_%04d.bi4" % k)):
            part = bi4.load(os.path.join(data, "Part_%04d.bi4" % k))
            it = part.items[0]
            idp = np.asarray(it.array("Idp")); pos = np.asarray(it.array("Posd")).reshape(-1, 3)
            nb = int(part.get("CaseNp") - part.get("CaseNfluid"))
            sel = idp < nb
            o = np.argsort(idp[sel])
            out["idp%d" % k] = idp[sel][o].astype(np.uint32); out["pos%d" % k] = pos[sel][o]
            out["time%d" % k] = it.get("TimeStep"); out["step%d" % k] = it.get("Step")
            k += 1
        out["nparts"] = k
        np.savez_compressed(os.path.join(dst, "bound.npz"), **out)
    json.dump({"spec": SPEC, "reference": "DualSPHysics v5.4 CPU (oracle/_ref)"}, open(os.path.join(dst, "fixture_motionref.json"), "w"), indent=1)
