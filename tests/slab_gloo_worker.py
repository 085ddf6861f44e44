"""Worker of tests/test_slab_host.py: runs the slab exchange protocol (dualsphysics_b200.slab._exchange_dist and the dt
all-reduce) over torch.distributed/gloo on CPU with a numpy stand-in for the CUDA context (1-D particles, records of
16 bytes: id u32, pad u32, x f64).  Launched with torch.distributed.run, world_size 2 or 3."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dualsphysics_b200 import slab  # noqa: E402

REC = np.dtype([("id", "<u4"), ("pad", "<u4"), ("x", "<f8")])


class FakeCtx:
    def slab_record_bytes(self, kind):
        return REC.itemsize


class FakeRank:
    """Same surface as slab.SlabRank (pack/unpack/sync/send_*/recv_*/ctx) over numpy arrays."""

    def __init__(self, rank, nranks, cuts, ids, x, scell):
        self.rank, self.nranks, self.torch, self.dev = rank, nranks, torch, torch.device("cpu")
        self.cxa, self.cxb, self.scell = cuts[rank], cuts[rank + 1], scell
        own = (np.floor(x / scell) >= self.cxa) & (np.floor(x / scell) < self.cxb)
        self.ids, self.x = ids[own].copy(), x[own].copy()
        self.ghost_ids = np.zeros(0, np.uint32)
        self.ctx = FakeCtx()
        mk = lambda: torch.zeros(len(ids) * REC.itemsize, dtype=torch.uint8)  # noqa: E731
        self.send_l, self.send_r, self.recv_l, self.recv_r = mk(), mk(), mk(), mk()

    def _put(self, buf, sel):
        r = np.zeros(int(sel.sum()), REC); r["id"], r["x"] = self.ids[sel], self.x[sel]
        buf[:r.nbytes] = torch.from_numpy(r.view(np.uint8).copy())
        return len(r)

    def pack(self, kind):
        cx = np.floor(self.x / self.scell)
        hl, hr = self.rank > 0, self.rank < self.nranks - 1
        if kind == slab.HALO:
            self.ghost_ids = np.zeros(0, np.uint32)
            l, r = (cx == self.cxa) & hl, (cx == self.cxb - 1) & hr
        else:
            l, r = (cx < self.cxa) & hl, (cx >= self.cxb) & hr
        n = self._put(self.send_l, l), self._put(self.send_r, r)
        if kind == slab.MIGRANT:
            keep = ~(l | r)
            self.ids, self.x = self.ids[keep], self.x[keep]
        return n

    def unpack(self, kind, nl, nr):
        for buf, n in ((self.recv_l, nl), (self.recv_r, nr)):
            if not n:
                continue
            r = buf[:n * REC.itemsize].numpy().view(REC)
            if kind == slab.HALO:
                self.ghost_ids = np.concatenate([self.ghost_ids, r["id"]])
            else:
                self.ids, self.x = np.concatenate([self.ids, r["id"]]), np.concatenate([self.x, r["x"]])

    def sync(self):
        pass

    def on_stream(self):
        import contextlib
        return contextlib.nullcontext()


def main():
    dist.init_process_group("gloo")
    rank, n = dist.get_rank(), dist.get_world_size()
    rng = np.random.default_rng(7)
    npart, ncx, scell = 4000, 40, 0.1
    x = rng.random(npart) ** 2 * (ncx * scell)           # denser towards x=0: the cuts must not be uniform
    vel = rng.normal(0, 0.04, npart)
    ids = np.arange(npart, dtype=np.uint32)
    cuts = slab.compute_cuts(np.floor(x / scell), ncx, n)
    fr = FakeRank(rank, n, cuts, ids, x, scell)
    for step in range(12):
        # "integrate": every rank moves its own particles (reflecting walls), then migrants + halos
        v = vel[fr.ids]
        fr.x = np.clip(fr.x + v, 1e-9, ncx * scell - 1e-9)
        x = np.clip(x + vel, 1e-9, ncx * scell - 1e-9)  # global truth replicated on every rank
        slab._exchange_dist(fr, slab.MIGRANT, dist)
        slab._exchange_dist(fr, slab.HALO, dist)
        cx = np.floor(x / scell)
        want = np.sort(ids[(cx >= fr.cxa) & (cx < fr.cxb)])
        assert np.array_equal(np.sort(fr.ids), want), f"rank {rank} step {step}: owned set wrong"
        assert np.allclose(fr.x[np.argsort(fr.ids)], x[want])
        wg = ids[((cx == fr.cxa - 1) & (rank > 0)) | ((cx == fr.cxb) & (rank < n - 1))]
        assert np.array_equal(np.sort(fr.ghost_ids), np.sort(wg)), f"rank {rank} step {step}: ghost set wrong"
        # dt maxima: ordered-uint MAX all-reduce, every rank must see the global maximum
        t = torch.tensor([int(np.abs(v).max() * 1e6) if len(v) else 0, rank, 5], dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        allv = [torch.zeros(3, dtype=torch.int32) for _ in range(n)]
        dist.all_gather(allv, t)
        assert all(torch.equal(a, t) for a in allv) and int(t[1]) == n - 1
    tot = torch.tensor([len(fr.ids)]); dist.all_reduce(tot)
    assert int(tot) == npart
    if rank == 0:
        print("SLAB_GLOO_OK", cuts)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
