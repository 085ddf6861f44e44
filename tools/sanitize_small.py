"""Small driver for compute-sanitizer (memcheck / racecheck): every kernel family on ~10k-particle cases, two steps each.

    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py [steps] [what,...]

what: dbc (variants 5, 0, 6; half and full cells; tile capacities that force the column split and the global fallback),
mdbc (warp-per-ghost-node kernels, SlipMode 1 and 2, laminar+SPS), float (floating box + piston), slab (two slabs in one
process through the library's device-counted exchange and a re-cut), 2d.  tests/test_gpu_sanitizer.py runs it under pytest."""
import os
import sys
import tempfile
from pathlib import Path

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import dualsphysics_b200 as dsb  # noqa: E402
from dualsphysics_b200 import slab  # noqa: E402
from tests import helpers as H  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
what = sys.argv[2].split(",") if len(sys.argv) > 2 else ["dbc", "mdbc", "float", "slab", "2d"]


def run(case, options, variant=5, cap=None):
    sim = dsb.Simulation(case, options=options)
    sim.ctx.set_variant(variant)
    if cap is not None:
        sim.ctx.set_tile_cap(cap)
    sim.ctx.run_steps(steps)
    info = sim.ctx.info()
    print(f"variant {variant} {options!r} cap {cap}: np {info.np} dt {info.dt_last:.6e}", flush=True)
    sim.close()


with tempfile.TemporaryDirectory() as d:
    d = Path(d)
    if "dbc" in what:
        H.write_case(d / "a", dp=0.025)
        for options in ("-cellmode:half", ""):
            run(str(d / "a" / "Case"), options, 5)
            run(str(d / "a" / "Case"), options, 0)
            for cap in (0, 150):
                run(str(d / "a" / "Case"), options, 6, cap)
    if "mdbc" in what:
        for slip in (1, 2):
            H.write_case(d / f"m{slip}", dp=0.04, layers=4, normals=True,
                         params={"Boundary": 2, "SlipMode": slip, "ViscoTreatment": 2, "Visco": 1e-6, "DensityDT": 3})
            run(str(d / f"m{slip}" / "Case"), "-cellmode:half", 5)
    if "float" in what:
        H.write_case(d / "f", dp=0.04, piston=True, piston_ampl=0.05, piston_freq=2.0, floating_box=(0.45, 0.25, 0.2, 0.1, 0.15, 0.1),
                     floating_rel_density=0.5)
        run(str(d / "f" / "Case"), "-cellmode:half", 5)
        run(str(d / "f" / "Case"), "-verlet -cubic", 5)
    if "2d" in what:
        H.write_case(d / "t", dp=0.01, two_d=True)
        run(str(d / "t" / "Case"), "", 5)
    if "slab" in what:
        H.write_case(d / "s", dp=0.03)
        ss = slab.SlabSimulation(str(d / "s" / "Case"), options="-cellmode:half", nranks=2, dist=None)
        assert ss.native, ss.cuts
        for _ in range(steps):
            ss.step()
        ss.recut()
        ss.step()
        print(f"slabs x2: cuts {ss.cuts} owned {ss.owned_counts()}", flush=True)
        ss.close()
print("sanitize_small done")
