#!/usr/bin/env python
"""A/B on the same box: the reference's OWN GPU solver (oracle/_ref/DualSPHysics5.4_linux64, unmodified sources compiled for
sm_100 with the reference's FAST_COMPILATION switch, oracle/Makefile `refgpu`) against this repository's path, on the bench block.

    python tools/ref_gpu_ab.py [np] [steps] > gpurun_out/TAG_ref_gpu_ab.json

The reference build only holds Wendland + artificial viscosity + DDT2Full (-ddt:3), so BOTH sides run -ddt:3 here (the bench line
itself stays on -ddt:2).  Per side: ms per step, and the interaction's share -- the reference from its own timers
(-svtimers:1: CF-Forces / NL-* / SU-ComputeStep in Run.out, src/source/JDsTimersGpu.h:74-96), ours from CUDA events.
Informational: it says how far the rebuilt kernel is from "the kernel to beat" (SURVEY 8c), it is not a parity check."""
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

REFGPU = os.path.join(ROOT, "oracle", "_ref", "DualSPHysics5.4_linux64")
BASE = "-symplectic -wendland -viscoart:0.1 -ddt:3 -ddtvalue:0.1 -dbc"


def run_ref(case, nsteps, cellmode):
    out = tempfile.mkdtemp(prefix="b2s_refgpu_")
    cmd = [REFGPU, case, out, "-gpu", "-sv:none", "-svtimers:1", f"-nsteps:{nsteps}", f"-cellmode:{cellmode}"] + BASE.split()
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    txt = r.stdout
    ro = os.path.join(out, "Run.out")
    if os.path.exists(ro):
        txt = open(ro, errors="replace").read()
    subprocess.run(["rm", "-rf", out])
    res = {"rc": r.returncode}
    m = re.search(r"Simulation Runtime\.*: *([0-9.eE+-]+) sec", txt)
    n = re.search(r"Steps of simulation\.*: *([0-9,]+)", txt)
    if not (m and n):
        res["error"] = txt[-1500:]
        return res
    res["steps"] = int(n.group(1).replace(",", ""))
    res["runtime_s"] = float(m.group(1))
    tm = {}
    for name, val in re.findall(r"^\s*((?:VA|NL|CF|SU)-[A-Za-z]+)\.*: *([0-9.eE+-]+) sec", txt, re.M):
        tm[name] = float(val)
    res["timers_s"] = tm
    return res


def main():
    npt = int(sys.argv[1]) if len(sys.argv) > 1 else 10000000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    case = bench.make_case(npt, "gpu")
    out = {"np_target": npt, "options": BASE, "reference_gpu": {}, "b200": {}}
    if not os.path.exists(REFGPU):
        out["reference_gpu"] = {"unavailable": "oracle/_ref/DualSPHysics5.4_linux64 not built (make -C oracle refgpu)"}
    else:
        for cm in ("full", "half"):
            a = run_ref(case, 3, cm)                   # warm-up run: its time is subtracted (the solver always starts at t=0)
            b = run_ref(case, 3 + steps, cm)
            if "error" in a or "error" in b:
                out["reference_gpu"][cm] = {"error": a.get("error") or b.get("error")}
                continue
            ns = b["steps"] - a["steps"]
            d = {k: (b["timers_s"].get(k, 0.0) - a["timers_s"].get(k, 0.0)) / ns * 1e3 for k in b["timers_s"]}
            out["reference_gpu"][cm] = {"steps": ns, "ms_per_step": (b["runtime_s"] - a["runtime_s"]) / ns * 1e3,
                                        "ms_per_step_by_timer": {k: round(v, 4) for k, v in d.items() if v > 1e-4},
                                        "ms_per_interaction": d.get("CF-Forces", 0.0) / 2}
    import torch
    import dualsphysics_b200 as dsb
    for cm in ("full", "half"):
        for variant in (0, 5):
            sim = dsb.Simulation(case, options=BASE + f" -cellmode:{cm}", device=0)
            ctx = sim.ctx
            ctx.set_variant(variant)
            ctx.run_steps(3)
            ks = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", 0))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(ks)
            ctx.run_steps(steps)
            e1.record(ks)
            torch.cuda.synchronize()
            out["b200"][f"{cm}_v{variant}"] = {"steps": steps, "ms_per_step": e0.elapsed_time(e1) / steps,
                                               "ms_per_interaction": ctx.time_interaction(5), "np": int(ctx.info().np)}
            sim.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
