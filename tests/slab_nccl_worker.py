"""Worker of tests/test_gpu_dist.py: one rank per GPU (torchrun, NCCL).  Runs a small dam break as slabs over the real
transport (torch.distributed batched isend/irecv + all-reduce on the ctx stream), gathers the owned particles on rank 0
and compares them with a single-context run of the same case.  usage: slab_nccl_worker.py <casepath> <options> <nsteps> [recut]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dualsphysics_b200 as dsb  # noqa: E402
from dualsphysics_b200 import slab  # noqa: E402


def main():
    case, options, nsteps = sys.argv[1], sys.argv[2], int(sys.argv[3])
    recut = len(sys.argv) > 4 and sys.argv[4] == "recut"
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ss = slab.SlabSimulation(case, options=options, nranks=world, dist=dist, device=lr, recut_every=(7 if recut else 0))
    dts = [ss.step() for _ in range(nsteps)]
    ids, pos, vr = ss.gather_byid()
    n = torch.tensor([len(ids)], device="cuda", dtype=torch.int64)
    counts = [torch.zeros(1, device="cuda", dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    nmax = max(counts)
    buf = torch.zeros(nmax, 8, dtype=torch.float64, device="cuda")
    buf[:len(ids), 0] = torch.from_numpy(ids.astype(np.float64)).cuda()
    buf[:len(ids), 1:4] = torch.from_numpy(pos).cuda()
    buf[:len(ids), 4:8] = torch.from_numpy(vr.astype(np.float64)).cuda()
    allb = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(allb, buf)
    migrated = torch.tensor([ss.migrated_total()], device="cuda", dtype=torch.int64)
    dist.all_reduce(migrated)
    if rank == 0:
        rows = np.concatenate([b[:c].cpu().numpy() for b, c in zip(allb, counts)])
        rows = rows[np.argsort(rows[:, 0], kind="stable")]
        ref = dsb.Simulation(case, options=options, device=lr)
        dts_ref = [ref.ctx.compute_step() for _ in range(nsteps)]
        rid = ref.ctx.array("idp"); o = np.argsort(rid, kind="stable")
        assert np.array_equal(rows[:, 0].astype(np.uint32), rid[o]), "particle set differs from the single-context run"
        epos = float(np.max(np.abs(rows[:, 1:4] - ref.ctx.array("pos")[o])))
        rv = ref.ctx.array("velrho")[o].astype(np.float64)
        evr = float(np.max(np.abs(rows[:, 4:8] - rv)) / np.max(np.abs(rv)))
        edt = max(abs(a - b) / b for a, b in zip(dts, dts_ref))
        print(f"SLAB_NCCL_RESULT world={world} cuts={ss.cuts} counts={counts} migrated={int(migrated.item())} pos_err={epos:.3e} "
              f"velrho_err={evr:.3e} dt_err={edt:.3e}")
        assert epos < 1e-6 and evr < 1e-4 and edt < 2e-5
        print("SLAB_NCCL_OK")
        ref.close()
    ss.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
