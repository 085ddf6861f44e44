#!/usr/bin/env python
"""Generates tests/golden/longrun_*.json: integrated observables of a LONG run of the compiled reference CPU solver
(oracle/_ref/DualSPHysics5.4CPU_linux64), against which tests/test_gpu_longrun.py checks this repository's driver.
north_star: "long runs must match on integrated observables (free-surface gauges, floating-body trajectory) within a stated
percentage, since SPH is chaotic".  Run here (needs /root/reference built by `make -C oracle ref`):  python tools/make_longrun_fixture.py"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from dualsphysics_b200 import bi4  # noqa: E402
from oracle import oracle as orc  # noqa: E402

REF_BIN = os.path.join(os.path.dirname(orc.REF_PROBE), "DualSPHysics5.4CPU_linux64")

CASES = {
    "longrun_dambreak_float": dict(
        spec=dict(dp=0.02, floating_box=(0.9, 0.235, 0.0, 0.12, 0.2, 0.12), floating_rel_density=0.6,
                  params={"TimeMax": 0.8, "TimeOut": 0.05}),
        options=["-symplectic", "-wendland", "-viscoart:0.1", "-ddt:2", "-ddtvalue:0.1", "-dbc"]),
}


def main():
    for name, cfg in CASES.items():
        with tempfile.TemporaryDirectory() as td:
            H.write_case(td, **cfg["spec"])
            outdir = os.path.join(td, "out")
            cmd = [REF_BIN, os.path.join(td, "Case"), outdir, "-cpu", "-sv:binx"] + cfg["options"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            assert r.returncode == 0 and "Finished execution (code=0)" in r.stdout, r.stdout[-3000:]
            obs = H.part_observables(os.path.join(outdir, "data"), cfg["spec"]["dp"])
        json.dump({"spec": cfg["spec"], "options": cfg["options"], "reference": "DualSPHysics v5.4 CPU (oracle/_ref), OpenMP default threads",
                   "observables": obs}, open(os.path.join(ROOT, "tests", "golden", name + ".json"), "w"), indent=1)
        print(name, len(obs), "parts; last:", obs[-1])


if __name__ == "__main__":
    main()
