"""Driver for ncu captures of the fluid interaction kernel: loads a dam-break block, warms up, launches the kernel alone."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import dualsphysics_b200 as dsb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 1
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 20
case = bench.make_case(n, "prof")
sim = dsb.Simulation(case, options=bench.OPTIONS)
sim.ctx.set_variant(variant)
sim.ctx.run_steps(warm)
ms = sim.ctx.time_interaction(2)
chk, real = sim.ctx.count_pairs()
print(f"np={sim.ctx.info().np} variant={variant} ms_per_launch={ms:.3f} checked={chk} real={real}")
sim.close()
