"""Multi-GPU run of a case with PART output, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        -m dualsphysics_b200.run Case_out/Case out_dir -symplectic -wendland -viscoart:0.1 -ddt:2 [-tmax:t] [-tout:t] [-nsteps:n]

Every rank loads its slab of the case (csrc/host/sim.cpp b2s_sim_load_slab), steps it with the halo / migrant exchange of slab.py
over NCCL, and writes its piece of every PART (Part_pRR_XXXX.bi4, the reference's piece format, JPartDataBi4.cpp:99-127) at the
times the single-GPU driver would (JDsOutputTime::GetNextTime, as in b2s_sim_run).  A single process without torchrun runs the
same loop on one GPU.  The switches are the reference's own; they go to the host driver unchanged.
"""
from __future__ import annotations

import os
import sys
import time
import xml.etree.ElementTree as ET


def next_part_time(t: float, tout: float) -> float:
    """First multiple of tout after t (the PART schedule of b2s_sim_run / JDsOutputTime::GetNextTime)."""
    ct = int(t / tout)
    tn = tout * ct
    while tn <= t:
        ct += 1
        tn = tout * ct
    return tn


def run_times(casepath: str, options: list) -> tuple:
    """(TimeMax, TimeOut, nsteps) from the case's <execution><parameters> with the -tmax / -tout / -nsteps overrides."""
    xml = casepath if casepath.endswith(".xml") else casepath + ".xml"
    prm = {e.get("key"): e.get("value") for e in ET.parse(xml).getroot().find("execution").find("parameters") if e.get("key")}
    tmax, tout, nsteps = float(prm.get("TimeMax", 0)), float(prm.get("TimeOut", 0)), 0
    for o in options:
        k, _, v = o.partition(":")
        k = k.lower()
        if k == "-tmax":
            tmax = float(v)
        elif k == "-tout":
            tout = float(v)
        elif k == "-nsteps":
            nsteps = int(v)
    if tmax <= 0 or tout <= 0:
        raise ValueError("TimeMax and TimeOut must be positive")
    return tmax, tout, nsteps


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) < 2:
        print(__doc__)
        return 1
    case, dirout, options = argv[0], argv[1], argv[2:]
    tmax, tout, nsteps = run_times(case, options)
    import torch
    from . import slab
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: this package has no CPU path")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ss = slab.SlabSimulation(case, options=" ".join(options), nranks=world, dist=dist, device=local)
    ss.set_dirout(dirout)
    if rank == 0 and world > 1:
        print(f"b200sph x{world}: slab step {'inside the library (NCCL from C++)' if ss.native else 'driven by slab.py (torch.distributed)'}; "
              "PART and PartOut files are written as pieces; PartMotionRef / PartExtra are not written in multi-GPU runs")
    ctx = ss.ranks[0].ctx
    try:                                     # a case with <special><gauges> registers them in the library: this runner has no CSV writer
        ctx.gauge_result(0)
        has_gauges = True
    except Exception:
        has_gauges = False
    if has_gauges:
        raise SystemExit("this case configures gauges: run it with the command-line driver (dualsphysics_b200/lib/b200sph, one process per GPU), "
                         "which evaluates them over the slabs and writes the CSV files; this runner would drop them")
    t = ctx.info().timestep
    ss.save_part()
    tnext, parts, steps, t0 = next_part_time(t, tout), 1, 0, time.time()
    while t < tmax:
        ss.step()
        steps += 1
        t = ctx.info().timestep
        if t >= tnext:
            ss.save_part()
            parts += 1
            tnext = next_part_time(t, tout)
        if nsteps and steps >= nsteps:
            break
    if dist is not None:
        dist.barrier()
    if rank == 0:
        secs = time.time() - t0
        print(f"b200sph x{world}: {steps} steps, simulated time {t:g} s, {parts} PART files ({world} pieces each), runtime {secs:.3f} s")
        print("\nFinished execution (code=0).")
    ss.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
