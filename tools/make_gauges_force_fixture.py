#!/usr/bin/env python
"""Generates tests/golden/gauges_force_ref/: GaugesForce_*.csv written by the UNMODIFIED reference CPU build
(oracle/_ref/DualSPHysics5.4CPU_linux64, -ompthreads:1) for a dam break with a piston: one force gauge on the fixed tank
(mkbound 0), one on the moving piston (mkbound 1), and a maxz gauge over the fluid column (GaugesMaxZ_Col.csv)."""
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dualsphysics_b200 import casegen  # noqa: E402

GAUGES = """<gauges>
  <default>
    <savevtkpart value="false" />
    <computedt value="0.002" />
    <computetime start="0" end="10" />
    <output value="true" />
    <outputdt value="0.002" />
    <outputtime start="0" end="10" />
  </default>
  <force name="Tank"><target mkbound="0" /></force>
  <force name="Piston"><target mkbound="1" /></force>
  <maxz name="Col"><point0 x="0.3" y="0.3" z="0.02" /><height value="0.5" /><distlimit coefdp="3" /></maxz>
</gauges>"""
SPEC = dict(dp=0.03, piston=True, piston_ampl=0.05, piston_freq=2.0, params={"TimeMax": 0.05, "TimeOut": 0.01}, special_xml=GAUGES)
OPTIONS = ["-symplectic"]

if __name__ == "__main__":
    dst = os.path.join(ROOT, "tests", "golden", "gauges_force_ref")
    os.makedirs(dst, exist_ok=True)
    with tempfile.TemporaryDirectory() as td:
        casegen.write_case(td, "Case", casegen.DamBreakSpec(**SPEC))
        exe = os.path.join(ROOT, "oracle", "_ref", "DualSPHysics5.4CPU_linux64")
        subprocess.run([exe, os.path.join(td, "Case"), os.path.join(td, "out"), "-cpu", "-ompthreads:1", "-sv:none"] + OPTIONS, check=True,
                       stdout=subprocess.DEVNULL)
        for f in ("GaugesForce_Tank.csv", "GaugesForce_Piston.csv", "GaugesMaxZ_Col.csv"):
            shutil.copy(os.path.join(td, "out", f), os.path.join(dst, f))
    json.dump({"spec": SPEC, "options": OPTIONS, "reference": "DualSPHysics v5.4 CPU (oracle/_ref), -ompthreads:1"},
              open(os.path.join(dst, "fixture_gauges.json"), "w"), indent=1)
