#!/usr/bin/env python
"""Generates tests/golden/wavetank_ref/: the files the UNMODIFIED reference CPU build writes for a 1 s wave-tank run that uses
the rows next to the hot path together -- piston wavemaker (mvrectsinu), floating box, standard shifting (NoFixed, threshold
1.5), a damping zone in front of the end wall, SWL / velocity / force gauges: GaugesSWL_*.csv, GaugesVel_*.csv,
GaugesForce_Piston.csv and PartFloatInfo.ibi4.  tests/test_gpu_longrun.py runs the same case through this repository's driver
and compares the time series.  Needs oracle/_ref (make -C oracle ref)."""
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dualsphysics_b200 import casegen  # noqa: E402

SPECIAL = """<gauges>
  <default>
    <savevtkpart value="false" />
    <computedt value="0.01" />
    <computetime start="0" end="10" />
    <output value="true" />
    <outputdt value="0.01" />
    <outputtime start="0" end="10" />
  </default>
  <swl name="Near"><pointdp coefdp="0.5" /><point0 x="0.5" y="0.34" z="0.04" /><point2 x="0.5" y="0.34" z="0.36" /></swl>
  <swl name="Far"><pointdp coefdp="0.5" /><point0 x="1.1" y="0.34" z="0.04" /><point2 x="1.1" y="0.34" z="0.36" /></swl>
  <velocity name="Mid"><point x="0.5" y="0.34" z="0.08" /></velocity>
  <force name="Piston"><target mkbound="1" /></force>
</gauges>
<damping>
  <dampingzone>
    <limitmin x="1.2" y="0" z="0" /><limitmax x="1.5" y="0" z="0" />
    <overlimit value="0.1" /><redumax value="10" /><factorxyz x="1" y="1" z="1" />
  </dampingzone>
</damping>"""
SPEC = dict(dp=0.02, fluid=(1.6, 0.67, 0.15), piston=True, piston_ampl=0.03, piston_freq=1.5,
            floating_box=(0.8, 0.24, 0.10, 0.12, 0.2, 0.08), floating_rel_density=0.5,
            params={"TimeMax": 1.0, "TimeOut": 0.1, "Shifting": 2, "ShiftCoef": -2, "ShiftTFS": 1.5}, special_xml=SPECIAL)
OPTIONS = ["-symplectic", "-wendland", "-viscoart:0.1", "-ddt:2", "-ddtvalue:0.1", "-dbc"]
FILES = ("GaugesSWL_Near.csv", "GaugesSWL_Far.csv", "GaugesVel_Mid.csv", "GaugesForce_Piston.csv")

if __name__ == "__main__":
    dst = os.path.join(ROOT, "tests", "golden", "wavetank_ref")
    os.makedirs(dst, exist_ok=True)
    with tempfile.TemporaryDirectory() as td:
        info = casegen.write_case(td, "Case", casegen.DamBreakSpec(**SPEC))
        exe = os.path.join(ROOT, "oracle", "_ref", "DualSPHysics5.4CPU_linux64")
        t0 = time.time()
        r = subprocess.run([exe, os.path.join(td, "Case"), os.path.join(td, "out"), "-cpu", "-sv:binx"] + OPTIONS, capture_output=True, text=True)
        assert r.returncode == 0 and "Finished execution (code=0)" in r.stdout, r.stdout[-3000:]
        secs = time.time() - t0
        for f in FILES:
            shutil.copy(os.path.join(td, "out", f), os.path.join(dst, f))
        shutil.copy(os.path.join(td, "out", "data", "PartFloatInfo.ibi4"), os.path.join(dst, "PartFloatInfo.ibi4"))
        steps = [ln for ln in r.stdout.splitlines() if "Steps of simulation" in ln or "Total Runtime" in ln or "Particles of simulation" in ln]
    json.dump({"spec": SPEC, "options": OPTIONS, "np": info["np"], "reference_seconds": secs, "reference_log": steps,
               "reference": "DualSPHysics v5.4 CPU (oracle/_ref), OpenMP default threads"},
              open(os.path.join(dst, "fixture_wavetank.json"), "w"), indent=1)
    print(info["np"], "particles,", secs, "s", steps)
