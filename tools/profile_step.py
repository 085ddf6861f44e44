#!/usr/bin/env python
"""Driver for `ncu --profile-from-start off` captures of ONE whole time step (every kernel of the step, not only the interaction):

    ncu --set full --clock-control none --import-source on --profile-from-start off \\
        --metrics sm__sass_thread_inst_executed_op_fadd_pred_on.sum,sm__sass_thread_inst_executed_op_fmul_pred_on.sum,sm__sass_thread_inst_executed_op_ffma_pred_on.sum \\
        -f -o gpurun_out/TAG_step python tools/profile_step.py bench 10000000 30

kind: bench (BASELINE configs[1]: DBC, artificial viscosity, DDT2), c3 (mDBC + laminar+SPS), c4 (piston + floating box).
The profiler is switched on around the last step only (cudaProfilerStart/Stop), after `warm` untimed steps."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402
import bench  # noqa: E402
import dualsphysics_b200 as dsb  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "bench"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10000000
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 30
if kind == "bench":
    case, options = bench.make_case(n, "prof"), bench.OPTIONS
else:
    import run_config
    case = run_config.make_case(kind, n, 1)
    options = "-symplectic -wendland -cellmode:half" + (" -viscoart:0.1 -ddt:2 -ddtvalue:0.1 -dbc" if kind == "c4" else "")
sim = dsb.Simulation(case, options=options)
sim.ctx.run_steps(warm)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
sim.ctx.run_steps(1)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(f"kind={kind} np={sim.ctx.info().np} profiled step {warm}")
sim.close()
