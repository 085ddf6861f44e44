#!/usr/bin/env python
"""bench.py -- WCSPH time-step throughput on B200 (metric of BASELINE.json: particle-steps/s, whole box).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's own CPU/OpenMP path

A "step" is one full ComputeStep of the synthetic 3-D dam-break block (BASELINE.json configs[1]: ~10M particles,
Symplectic, Wendland, artificial viscosity 0.1, Fourtakas DDT, DBC), i.e. 2 cell divides + 2 interactions +
predictor/corrector.  Prints ONE JSON line.  Keys beyond the base contract:
  roofline      dominant kernel (fluid interaction) against its binding roof (FP32 issue)
                + "hbm" views: the same kernel's algorithmic HBM bytes and the whole step's 0.71 KB/particle-step
  cpu_baseline  the compiled reference CPU solver (oracle/_ref) on this box's host cores, bounded sample
  e2e           same metric through the public API with HOST buffers: per step upload (pinned H2D) + step + download
"""
import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-steps/sec (whole box) at 1/2/4/8 B200; interaction-kernel roofline fraction"
UNIT = "particle-steps/s"
OPTIONS = "-symplectic -wendland -viscoart:0.1 -ddt:2 -ddtvalue:0.1 -dbc -cellmode:half" + os.environ.get("B2S_BENCH_EXTRA_OPTS", "")
# DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of ONE launch of the fluid interaction kernel at the 10M-particle
# workload, from the `ncu --set full` captures summarised under profiles/ (file, read MB + write MB); scaled by the fluid count
# when the bench runs another size.  A variant without a capture reports traffic null instead of a stale number.
NCU_TRAFFIC = {5: ("profiles/r02K_ncu_forces_pair2_v5_10M.txt", 459204352 + 180592128),
               6: ("profiles/r02i_ncu_forces_tile_v6_10M.txt", None)}
NCU_TRAFFIC_NPF = 9281220
# Non-tensor FP32 peak: tools/microbench.cu measures 3.798 FFMA warp-instructions per clock and SM of the nominal 4
# (profiles/r02b_microbench.txt), i.e. 148 SM x 128 lanes x 2 flop x clock x 0.9495
FP32_MEASURED_FRACTION = 3.798 / 4.0
NP_TARGET = int(os.environ.get("B2S_BENCH_NP", "10000000"))
CPU_SAMPLE_NP = int(os.environ.get("B2S_BENCH_CPU_NP", "1000000"))
PREROLL = int(os.environ.get("B2S_BENCH_PREROLL", "200"))
# flop model of SURVEY 8(d): per checked pair 16 flop, per real pair (r<2h) 90 flop
FLOP_CHK, FLOP_REAL = 16.0, 90.0
# algorithmic HBM bytes: interaction pass 50 B per fluid particle; whole Symplectic step 0.71 KB per particle
BYTES_INTER, BYTES_STEP = 50.0, 710.0


def peaks():
    f = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(f):
        d = json.load(open(f))
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured"
    return 6650.0, 1965.0, "fallback"


def make_case(np_target, tag, xscale=1):
    """The dam-break block at ~np_target particles per unit of xscale (tank and column stretched along x for weak scaling)."""
    from dualsphysics_b200 import casegen
    d = os.path.join(tempfile.gettempdir(), f"b2s_bench_{tag}_{np_target}_x{xscale}")
    if not os.path.exists(os.path.join(d, "Case.done")):
        dp = casegen.dp_for_target(np_target)
        casegen.write_case(d, "Case", casegen.DamBreakSpec(dp=dp, tank=(1.6 * xscale, 0.67, 0.4), fluid=(0.4 * xscale, 0.67, 0.3),
                                                           params={"TimeMax": 10.0}))
        open(os.path.join(d, "Case.done"), "w").write("ok")
    return os.path.join(d, "Case")


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if "Active" in r[5 + k] and "Not" not in r[5 + k]:
                        reasons.add(nm)
            except Exception:
                pass
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def run_reference_stock(np_target, steps, warmup, threads=None, tag="cpu"):
    """Times the reference's own stock CPU/OpenMP solver (oracle/_ref/DualSPHysics5.4CPU_linux64, compiled from the unmodified
    sources by oracle/Makefile) the way BASELINE.md prescribes: `-nsteps:K -sv:none -svtimers:1`, "Simulation Runtime" and
    "Steps of simulation" from Run.out (src/source/JSph.cpp:3534-3542, JSphCfgRun.cpp:243).  The solver always starts at t=0, so
    the warm-up steps are a second, shorter run whose simulation time is subtracted."""
    from oracle import oracle as orc
    if not os.path.exists(orc.REF_SOLVER):
        return None
    case = make_case(np_target, tag)
    ncores = os.cpu_count() or 1
    th = min(threads or ncores, 64)          # OMP_MAXTHREADS 64 (src/source/OmpDefs.h:39)

    def once(n):
        out = tempfile.mkdtemp(prefix="b2s_ref_")
        cmd = [orc.REF_SOLVER, case, out, "-cpu", f"-ompthreads:{th}", "-sv:none", "-svtimers:1", f"-nsteps:{n}"] + OPTIONS.split()
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        txt = r.stdout
        runout = os.path.join(out, "Run.out")
        if os.path.exists(runout):
            txt = open(runout, errors="replace").read()
        m1 = re.search(r"Simulation Runtime\.*: *([0-9.eE+-]+) sec", txt)
        m2 = re.search(r"Steps of simulation\.*: *([0-9,]+)", txt)
        m3 = re.search(r"Particles of simulation \(initial\): *([0-9,]+)", txt) or re.search(r"\bNp=([0-9]+)", txt)
        if not (m1 and m2):
            raise RuntimeError("reference solver failed: " + txt[-1500:])
        subprocess.run(["rm", "-rf", out])
        return int(m2.group(1).replace(",", "")), float(m1.group(1)), (int(m3.group(1).replace(",", "")) if m3 else 0)
    s_w, t_w, npc = once(warmup) if warmup else (0, 0.0, 0)
    s_a, t_a, npc = once(warmup + steps)
    if not npc:
        from dualsphysics_b200 import bi4
        npc = int(bi4.read_case_counts(case + ".bi4")) if hasattr(bi4, "read_case_counts") else np_target
    nst = max(s_a - s_w, 1)
    secs = max(t_a - t_w, 1e-9)
    return {"value": npc * nst / secs, "np": npc, "steps": nst, "seconds": secs, "cores": th, "ms_per_step": secs / nst * 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--np", type=int, default=NP_TARGET)
    ap.add_argument("--variant", type=int, default=int(os.environ.get("B2S_VARIANT", "5")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the variant-0 cross-check on the bench state")
    ap.add_argument("--no-pipeline", action="store_true", help="e2e with one block only (no double buffering)")
    a = ap.parse_args()
    # the contract is ONE JSON line on stdout: libraries that write to fd 1 (NCCL prints its version there) go to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(obj), flush=True)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(a.warmup, 3) if a.impl == "b200" else a.warmup

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if a.impl == "reference":
        if rank != 0:
            return 0
        # Same block as the GPU arm (a.np particles) when the host has the cores to finish K steps in minutes (~2M
        # particle-steps/s on 16 threads); otherwise, or with B2S_BENCH_CPU_NP_REF, a smaller block of the same shape.
        ncores = os.cpu_count() or 1
        np_ref = int(os.environ.get("B2S_BENCH_CPU_NP_REF", str(a.np if ncores >= 12 else CPU_SAMPLE_NP)))
        st, wu = max(1, a.steps), max(0, min(a.warmup, 5))
        res = run_reference_stock(np_ref, st, wu, tag=("gpu" if np_ref == a.np else "cpu"))
        if res is None:
            emit({"impl": "reference", "unavailable": "oracle/_ref was not built (needs /root/reference in the build container)"})
            return 0
        same = (np_ref == a.np)
        sample = (f"{res['np']} particles x {res['steps']} steps (after {wu} warm-up steps, subtracted) of the dam-break block"
                  + (" = the GPU arm's block" if same else " (smaller block of the same shape)") + f", lattice start, stock DualSPHysics5.4CPU_linux64 "
                  f"-cpu -ompthreads:{res['cores']} -nsteps -sv:none -svtimers:1, Simulation Runtime from Run.out")
        line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": res["steps"],
                "warmup": wu, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"synthetic 3D dam-break block (BASELINE configs[1]), {OPTIONS}; the reference's stock CPU/OpenMP solver",
                           "np": res["np"], "same_block_as_gpu_arm": same},
                "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": "reference", "sample": sample},
                "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        emit(line)
        return 0

    # ------------------------------------------------------------------ this repo's arm
    import torch
    import dualsphysics_b200 as dsb
    from dualsphysics_b200 import slab
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no fallback")
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank if world > 1 else 0
    torch.cuda.set_device(dev)
    hbm_gbs, sm_max_mhz, peak_src = peaks()

    # weak scaling: the block (tank and water column) is `world` times longer along x, ~a.np particles per GPU
    if rank == 0:
        case = make_case(a.np, "gpu", xscale=world)
    if dist is not None:
        dist.barrier()
    case = make_case(a.np, "gpu", xscale=world)
    ss = None
    if world == 1:
        sim = dsb.Simulation(case, options=OPTIONS, device=dev)
        ctx = sim.ctx
    else:
        ss = slab.SlabSimulation(case, options=OPTIONS, nranks=world, dist=dist, device=dev,
                                 recut_every=int(os.environ.get("B2S_RECUT_EVERY", "50")))   # dynamic load balance on measured time (csrc/slab_comm.cu grp_recut)
        sim = ss.ranks[0].sim
        ctx = ss.ranks[0].ctx
    ctx.set_variant(a.variant)
    info0 = ctx.info()
    if ss is None:
        np_total, npf = info0.np, info0.np - info0.npb
    else:
        code = ctx.array("code")
        own = (code & 0xe000) == 0
        np_total, npf = int(own.sum()), int((own & ((code & 0x1800) == 0x1800)).sum())

    def run(n):
        if ss is None:
            ctx.run_steps(n)
        else:
            ss.run_steps(n)       # the whole slab step runs inside the library (csrc/slab_comm.cu), one host sync per step

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- untimed pre-roll: the lattice start disorders (200 steps), so that the timed steps see the memory access
    #      pattern of a running simulation, not of a crystal
    run(PREROLL)
    # ---- warm-up (the clock sampler starts here: the warm-up is the same workload and nvidia-smi needs ~100 ms per sample)
    sampler = ClockSampler(dev) if rank == 0 else None
    run(warmup)
    # ---- timed steps: CUDA events on the stream the kernels are launched on (the ctx stream), bracketed by
    #      barrier + synchronize; max over ranks.  Wall clock is taken too and must agree.
    kstream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", dev))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.launch_count(reset=True)
    if ss is not None:
        ss.exchanged_bytes = 0
        ss.comm_stats(reset=True)
    barrier()
    t0 = time.perf_counter()
    ev0.record(kstream)
    run(a.steps)
    ev1.record(kstream)
    barrier()
    wall = time.perf_counter() - t0
    secs = ev0.elapsed_time(ev1) * 1e-3
    launches = ctx.launch_count()
    xbytes = ss.exchanged_bytes if ss is not None else 0
    comm = None
    if ss is not None and ss.native:
        ms_x, nx, bytes_x, nrecut, _ = ss.comm_stats()
        xbytes = bytes_x
        comm = [ms_x / a.steps, nx / a.steps, float(nrecut)]
        rep = ss.profile_report()       # B2S_SLAB_PROFILE=1: ms per step and phase of the timed steps, every rank
        if rep:
            reps = [None] * world
            dist.all_gather_object(reps, f"rank {rank} (np {np_total}, fluid {npf}): {rep}")
            if rank == 0:
                print("slab phase profile, timed steps:\n  " + "\n  ".join(reps), file=sys.stderr)
    if dist is not None:
        t = torch.tensor([secs, wall], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs, wall = float(t[0].item()), float(t[1].item())
        if comm is not None:
            tc = torch.tensor(comm, device="cuda", dtype=torch.float64)
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
            comm = [float(v) for v in tc.tolist()]
        cnt = torch.tensor([float(np_total), float(launches), float(xbytes)], device="cuda", dtype=torch.float64)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        np_all, launches_all, xbytes_all = int(cnt[0].item()), int(cnt[1].item()), int(cnt[2].item())
        allnp = [torch.zeros(1, device="cuda", dtype=torch.float64) for _ in range(world)]
        dist.all_gather(allnp, torch.tensor([float(np_total)], device="cuda", dtype=torch.float64))
        np_per_rank = [int(x.item()) for x in allnp]
    else:
        np_all, launches_all, xbytes_all, np_per_rank = np_total, launches, 0, [np_total]
    value = np_all * a.steps / secs
    clocks = sampler.stop() if sampler else None

    # ---- dominant kernel alone: CUDA events on the ctx stream (b2s_time_interaction), flop model from pair counts
    ms_k = ctx.time_interaction(5)
    chk, real = ctx.count_pairs()
    # ---- parity on the bench state itself (outside every timed region): the benched variant against variant 0 (reference
    #      summation order, the kernel every golden fixture pins) on the SAME particle state; max difference relative to max|field|
    parity = None
    if ss is None and a.variant != 0 and not a.no_parity:
        import numpy as np
        inter = dsb.lib.INTER_SYMPREDICTOR
        ctx.interaction(inter)
        fast = ctx.array("acear").copy()
        ctx.set_variant(0)
        ctx.interaction(inter)
        ref0 = ctx.array("acear").copy()
        ctx.set_variant(a.variant)
        own = (ctx.array("code") & 0xf800) == 0x1800          # fluid, not a ghost
        sc_a = float(np.abs(ref0[own, :3]).max()) or 1.0
        p_ = sim.params
        sc_r = max(float(np.abs(ref0[own, 3]).max()), float(p_.rhopzero) * p_.cs0 / float(p_.kernelh) * 1e-4)
        parity = {"against": "variant 0 on the same state", "particles": int(own.sum()),
                  "ace_max_rel": float(np.abs(fast[own, :3] - ref0[own, :3]).max()) / sc_a,
                  "ar_max_rel": float(np.abs(fast[own, 3] - ref0[own, 3]).max()) / sc_r}
        del fast, ref0
    # ---- e2e: host buffers in, host buffers out, every step (pinned memory; copies inside the timed region)
    e2e = None
    if not a.no_e2e:
        import ctypes as C
        n = ctx.info().np
        cap = int(n * 1.3) + 4096
        h_idp = torch.empty(cap, dtype=torch.int32).pin_memory()
        h_code = torch.empty(cap, dtype=torch.int16).pin_memory()
        h_pos = torch.empty(cap * 3, dtype=torch.float64).pin_memory()
        h_vr = torch.empty(cap * 4, dtype=torch.float32).pin_memory()
        lib = dsb.load_library()

        def download():
            cnt_ = C.c_uint32(0)
            if ss is not None:
                ss.host_roundtrip_begin()
                assert ctx.info().np <= cap
            r = lib.b2s_download_particles(ctx._h, C.byref(cnt_), h_idp.data_ptr(), h_code.data_ptr(), h_pos.data_ptr(), h_vr.data_ptr())
            assert r == 0, r
            return cnt_.value

        def upload(n_):
            ctx.upload_ptrs(n_, h_idp.data_ptr(), h_code.data_ptr(), h_pos.data_ptr(), h_vr.data_ptr())
            if ss is not None:
                ss.host_roundtrip_end()
        if ss is not None:
            ss.step(final_exchange=False)
        n = download()
        ksteps = max(2, min(a.steps, 5))
        moved = 0
        barrier()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            upload(n)
            moved += n
            if ss is None:
                ctx.compute_step()
            else:
                ss.step(final_exchange=False)
            n = download()
        barrier()
        se = time.perf_counter() - t0
        tot = torch.tensor([float(moved), se], device="cuda", dtype=torch.float64)
        if dist is not None:
            mx = tot.clone()
            dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            se = float(mx[1].item())
        moved_all = int(tot[0].item())
        bytes_pp = 4 + 2 + 24 + 16
        note = ("per step and per rank: b2s_upload_particles (pinned host -> device, incl. cell divide)"
                + ("" if ss is None else " + migrant/halo exchange") + " + one full step + b2s_download_particles (all owned particles)")
        e2e = {"value": moved_all / se, "unit": UNIT, "h2d_bytes_per_step": moved_all // ksteps * bytes_pp,
               "d2h_bytes_per_step": moved_all // ksteps * bytes_pp, "steps": ksteps, "pipeline_depth": 1, "note": note}
        if ss is None and not a.no_pipeline:
            # Double buffering: TWO independent blocks in flight, each on its own context/stream and its own pinned buffers,
            # driven by two host threads, so that the PCIe copies of one block overlap the kernels of the other.  Every step of
            # every block still uploads its inputs, runs the full step and downloads the result.
            import threading
            depth = max(2, int(os.environ.get("B2S_E2E_DEPTH", "3")))   # measured 2: 278M, 3: 301M, 4: 309M particle-steps/s
            sims = [dsb.Simulation(case, options=OPTIONS, device=dev) for _ in range(depth - 1)]
            for s_ in sims:
                s_.ctx.set_variant(a.variant)
                s_.ctx.run_steps(warmup)
            ctxs = [ctx] + [s_.ctx for s_ in sims]
            bufs = [(h_idp, h_code, h_pos, h_vr)] + [
                (torch.empty(cap, dtype=torch.int32).pin_memory(), torch.empty(cap, dtype=torch.int16).pin_memory(),
                 torch.empty(cap * 3, dtype=torch.float64).pin_memory(), torch.empty(cap * 4, dtype=torch.float32).pin_memory())
                for _ in sims]
            done = [0] * depth

            def worker(k, c, bufs_):
                cnt_ = C.c_uint32(0)
                ptrs = [b.data_ptr() for b in bufs_]
                assert lib.b2s_download_particles(c._h, C.byref(cnt_), *ptrs) == 0
                start.wait()
                for _ in range(ksteps):
                    c.upload_ptrs(cnt_.value, *ptrs)
                    done[k] += cnt_.value
                    c.compute_step()
                    assert lib.b2s_download_particles(c._h, C.byref(cnt_), *ptrs) == 0
            start = threading.Barrier(depth + 1)
            th = [threading.Thread(target=worker, args=(k, ctxs[k], bufs[k])) for k in range(depth)]
            for t_ in th:
                t_.start()
            torch.cuda.synchronize()
            start.wait()
            t0 = time.perf_counter()
            for t_ in th:
                t_.join()
            torch.cuda.synchronize()
            sp = time.perf_counter() - t0
            for s_ in sims:
                s_.close()
            # headline = ONE block (upload, step, download strictly in sequence); the figure with several blocks in flight is secondary
            e2e["pipelined"] = {"value": sum(done) / sp, "blocks_in_flight": depth,
                                "note": f"{depth} independent blocks (one context, host thread and set of pinned buffers each): the copies of "
                                        "one block overlap the kernels of the others; every step of every block still uploads, steps and downloads"}

    line = None
    if rank == 0:
        flops = FLOP_CHK * chk + FLOP_REAL * real
        tf_ach = flops / (ms_k * 1e-3) / 1e12
        fp32_peak = 148 * 128 * 2 * sm_max_mhz * 1e6 / 1e12 * FP32_MEASURED_FRACTION
        tr_file, tr_bytes = NCU_TRAFFIC.get(a.variant, (None, None))
        gbs_k = BYTES_INTER * npf / (ms_k * 1e-3) / 1e9
        gbs_step = BYTES_STEP * np_total / (secs / a.steps) / 1e9
        roofline = {"kernel": ({4: "k_forces_fluid_pair2", 5: "k_forces_fluid_pair2", 6: "k_forces_fluid_tile"}.get(a.variant, "k_forces_fluid_gather") + f" (fluid<-fluid+bound interaction, variant {a.variant})"), "bound": "fp32", "achieved": tf_ach, "peak": fp32_peak,
                    "unit": "TFLOP/s", "frac": tf_ach / fp32_peak, "traffic": (int(tr_bytes * npf / NCU_TRAFFIC_NPF) if tr_bytes else None),
                    "traffic_note": f"dram__bytes_read.sum+dram__bytes_write.sum per launch from {tr_file} (ncu --set full), scaled by fluid count; algorithmic bytes = 50 B x fluid particles",
                    "peak_source": "measured: 148 SM x 128 FP32 lanes x 2 x clocks.max.sm x 0.9495 = the FFMA issue rate tools/microbench.cu reaches on this part (profiles/r02b_microbench.txt); non-tensor FP32, tensor cores do not apply to pairwise SPH",
                    "ms_per_launch": ms_k, "pairs_checked": chk, "pairs_real": real, "flop_model": "16*checked + 90*real (SURVEY 8d)",
                    "share_of_step": 2 * ms_k / (secs / a.steps * 1e3),
                    "hbm": {"bound": "hbm", "achieved": gbs_k, "peak": hbm_gbs, "unit": "GB/s", "frac": gbs_k / hbm_gbs,
                            "algorithmic_bytes_per_particle": BYTES_INTER, "peak_source": peak_src},
                    "hbm_whole_step": {"bound": "hbm", "achieved": gbs_step, "peak": hbm_gbs, "unit": "GB/s", "frac": gbs_step / hbm_gbs,
                                       "algorithmic_bytes_per_particle_step": BYTES_STEP, "peak_source": peak_src}}
        # ---- CPU baseline: the reference's own CPU solver on this box's cores, bounded sample
        cpu = None
        if not a.no_cpu_baseline and world == 1:
            try:
                res = run_reference_stock(CPU_SAMPLE_NP, 5, 1)
                if res:
                    cpu = {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": "reference",
                           "sample": f"{res['np']} particles x {res['steps']} steps (1 warm-up step subtracted) of a smaller block of the same shape, stock oracle/_ref/DualSPHysics5.4CPU_linux64 (unmodified DualSPHysics v5.4 sources) -cpu -nsteps -sv:none, Simulation Runtime from Run.out, {res['cores']} OpenMP threads; the full block is timed by `bench.py --impl reference`"}
            except Exception as ex:  # noqa: BLE001
                cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"failed: {ex}"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": warmup,
                "ms_per_step": secs / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"synthetic 3D dam-break block (BASELINE configs[1]), {OPTIONS}, lattice start + {PREROLL} untimed steps"
                                       + ("" if world == 1 else f", block {world}x longer along x"),
                           "np_per_gpu": np_per_rank, "np_total": np_all, "fluid_rank0": npf, "interaction_variant": a.variant,
                           "parallelism": "single GPU" if world == 1 else f"{world} slabs along x (cuts balance work: fluid particle = 3 boundary particles), one rank per GPU; per interaction: migrants+halos in one fixed-size message per neighbour over ncclSend/ncclRecv issued by the library (counts stay on the device), dt maxima by ncclAllReduce(MAX); one host sync per step",
                           "comm": (None if comm is None else {"t_exposed_comm_ms_per_step": comm[0], "exchanges_per_step": comm[1], "recuts": comm[2],
                                                              "note": "max over ranks of the CUDA-event time inside pack + ncclSend/ncclRecv + unpack on the kernel stream; the exchange is serial with the kernels (not overlapped), so this whole time is exposed; it includes waiting for a slower neighbour"}),
                           "exchanged_bytes_per_step": xbytes_all // max(a.steps, 1),
                           "l2_note": "inputs larger than L2: working set ~2.3 GB per GPU >> 126 MB L2",
                           "timing": "CUDA events on the kernel stream, max over ranks", "wall_ms_per_step": wall / a.steps * 1e3},
                "roofline": roofline, "parity": parity, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches_all, "clocks": clocks}
    if ss is not None:
        ss.close()
    else:
        sim.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(line)
    return 0


if __name__ == "__main__":
    sys.exit(main())
