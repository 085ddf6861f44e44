#!/usr/bin/env python
"""Runs one of the BASELINE.json parity configurations at scale and prints ONE JSON line (evidence, not the bench contract):

    c3  dam break with mDBC boundaries (4 layers + normals, SlipMode 1) and laminar+SPS viscosity, DDT3
    c4  wave tank: piston wavemaker (mvrectsinu) + floating box + shallow water over the whole tank, DBC, artificial viscosity, DDT2

    python tools/run_config.py c3 --npgpu 5000000 --steps 10                       # one GPU
    python -m torch.distributed.run --nproc-per-node 2 ... tools/run_config.py c4 --npgpu 5000000   # slabs along x, tank N times longer

Timed like bench.py (CUDA events on the ctx stream, barrier + synchronize on both sides, max over ranks).  Checks on the way:
no particle is lost or excluded, dt stays inside [dtmin, dtini], the floating body stays finite and inside the tank."""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_case(cfg, np_target, world):
    from dualsphysics_b200 import casegen
    d = os.path.join(tempfile.gettempdir(), f"b2s_cfg_{cfg}_{np_target}_x{world}")
    if not os.path.exists(os.path.join(d, "Case.done")):
        if cfg == "c3":
            dp = casegen.dp_for_target(np_target, layers=4)
            spec = casegen.DamBreakSpec(dp=dp, tank=(1.6 * world, 0.67, 0.4), fluid=(0.4 * world, 0.67, 0.3), layers=4, normals=True,
                                        params={"TimeMax": 10.0, "Boundary": 2, "SlipMode": 1, "ViscoTreatment": 2, "Visco": 1e-6, "DensityDT": 3})
        else:
            dp = casegen.dp_for_target(np_target, fluid=(1.6, 0.67, 0.15))
            nd = lambda v: round(v / dp) * dp   # noqa: E731  (box on the lattice)
            spec = casegen.DamBreakSpec(dp=dp, tank=(1.6 * world, 0.67, 0.4), fluid=(1.6 * world, 0.67, 0.15), piston=True, piston_ampl=0.02, piston_freq=1.0,
                                        floating_box=(nd(0.8 * world), nd(0.235), nd(0.10), nd(0.12), nd(0.2), nd(0.08)), floating_rel_density=0.5,
                                        params={"TimeMax": 10.0})
        casegen.write_case(d, "Case", spec)
        open(os.path.join(d, "Case.done"), "w").write("ok")
    return os.path.join(d, "Case")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["c3", "c4"])
    ap.add_argument("--npgpu", type=int, default=5000000, help="particles per GPU (about)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    rank, world, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import numpy as np
    import torch
    import dualsphysics_b200 as dsb
    from dualsphysics_b200 import slab
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank if world > 1 else 0
    torch.cuda.set_device(dev)
    options = "-symplectic -wendland -cellmode:half" + (" -viscoart:0.1 -ddt:2 -ddtvalue:0.1 -dbc" if a.config == "c4" else "")
    t_gen = time.perf_counter()
    if rank == 0:
        make_case(a.config, a.npgpu, world)
    if dist is not None:
        dist.barrier()
    case = make_case(a.config, a.npgpu, world)
    t_gen = time.perf_counter() - t_gen
    ss = None
    if world == 1:
        sim = dsb.Simulation(case, options=options, device=dev)
        ctx = sim.ctx
    else:
        ss = slab.SlabSimulation(case, options=options, nranks=world, dist=dist, device=dev)
        sim, ctx = ss.ranks[0].sim, ss.ranks[0].ctx
    p = sim.params

    def owned():
        code = ctx.array("code")
        return int(((code & 0xe000) == 0).sum())

    def total(v):
        if dist is None:
            return v
        t = torch.tensor([float(v)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())
    n0 = total(owned())
    dts = []

    def run(n):
        for _ in range(n):
            dts.append(ctx.compute_step() if ss is None else ss.step())

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
    run(a.warmup)
    kstream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", dev))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(kstream)
    run(a.steps)
    ev1.record(kstream)
    barrier()
    secs = ev0.elapsed_time(ev1) * 1e-3
    if dist is not None:
        t = torch.tensor([secs], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs = float(t.item())
    n1 = total(owned())
    info = ctx.info()
    comm = None
    if ss is not None and ss.native:
        ms_x, nx, bytes_x, _, _ = ss.comm_stats()
        comm = {"t_exposed_comm_ms_per_step_rank0": ms_x / max(len(dts), 1), "exchanges": nx, "bytes_sent_rank0": bytes_x, "phases_ms_per_step_rank0": ss.profile_report()}
    out = {"slab_step": (None if ss is None else ("library (csrc/slab_comm.cu, NCCL)" if ss.native else "slab.py (torch.distributed)")), "comm": comm,
           "interaction_variant": 5, "config": a.config, "n_gpus": world, "np_total": n0, "np_after": n1, "casenp": int(p.casenp), "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": secs / a.steps * 1e3, "particle_steps_per_s": n0 * a.steps / secs, "options": options,
           "dt_first": dts[0], "dt_last": dts[-1], "dt_min": min(dts), "timestep": info.timestep, "case_seconds": t_gen,
           "cells": [int(info.ncx), int(info.ncy), int(info.ncz)]}
    ok = (n0 == int(p.casenp) and n1 == n0 and all(np.isfinite(dts)) and min(dts) >= p.dtmin * 0.999)
    if a.config == "c4":
        ft = ctx.get_floatings(1)[0]
        cen = [ft.center[k] for k in range(3)]
        out["float_center"] = cen
        out["float_vel"] = [ft.fvel[k] for k in range(3)]
        ok = ok and all(np.isfinite(cen)) and 0 < cen[2] < 0.4
    out["ok"] = bool(ok)
    if rank == 0:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(out), flush=True)
    if ss is not None:
        ss.close()
    else:
        sim.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
