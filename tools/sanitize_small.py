"""Small driver for compute-sanitizer (memcheck / racecheck): a 20k-particle dam break stepped with interaction variant 6 at the
default tile capacity and at capacities that force the column split and the global-memory fallback; DBC and mDBC+SPS-free
configurations the tile kernel covers.  Usage: compute-sanitizer --tool memcheck python tools/sanitize_small.py [steps]"""
import sys, os, tempfile
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import dualsphysics_b200 as dsb
from tests import helpers as H
from pathlib import Path

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [6]
with tempfile.TemporaryDirectory() as d:
    H.write_case(Path(d), dp=0.02)
    for variant in variants:
        for options in ("-cellmode:half", ""):
            for cap in (0, 700, 150):
                sim = dsb.Simulation(os.path.join(d, "Case"), options=options)
                sim.ctx.set_variant(variant)
                sim.ctx.set_tile_cap(cap)
                sim.ctx.run_steps(steps)
                info = sim.ctx.info()
                print(f"variant {variant} {options!r} cap {cap}: np {info.np} dt {info.dt_last:.6e}")
                sim.close()
print("sanitize_small done")
