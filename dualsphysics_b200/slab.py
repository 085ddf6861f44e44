"""Slab-decomposed multi-GPU driver: one process (rank) per GPU, NCCL send/recv over NVLink for halos and migrants,
NCCL all-reduce(MAX) for the global dt.  No reference counterpart (DualSPHysics v5.4 is single-GPU).

Design:
  * the domain is cut at cell planes along x; cuts come from the prefix sum of particles per x cell plane so that every
    rank owns ~Np/G particles (``compute_cuts``);
  * per interaction: migrants (particles that left the slab, full state) then halos (one cell layer each side, appended
    as CODE_PERIODIC ghosts) are packed into contiguous device buffers by the CUDA side (csrc/slab.cu) and exchanged with
    the two neighbours; then the usual cell divide + interaction run on owned+ghost particles;
  * the three dt maxima are all-reduced with MAX so that every rank integrates with the same dt.
The transport is pluggable: ``TorchDistTransport`` (NCCL, one process per GPU) or ``LocalTransport`` (all ranks in one
process on one GPU -- used by the tests to check the protocol against the single-context result).
"""
from __future__ import annotations

import numpy as np

from . import lib as L

HALO, MIGRANT = 0, 1


def compute_cuts(cellx: np.ndarray, ncells_x: int, nranks: int, minwidth: int = 1) -> list:
    """Cut planes [c0=0, c1, ..., cN=ncells_x] balancing the particle count per slab (prefix sum over x cell planes)."""
    return compute_cuts_hist(np.bincount(np.asarray(cellx, np.int64), minlength=ncells_x), nranks, minwidth)


def compute_cuts_hist(hist, nranks: int, minwidth: int = 1) -> list:
    """Same from the histogram of particles per x cell plane (the form the re-cut uses after an all-reduce).
    ``minwidth``: planes per slab at least (the halo is scelldiv planes wide and only comes from the direct neighbours, so a
    slab must be at least that wide)."""
    hist = np.asarray(hist, np.int64)
    ncells_x = len(hist)
    cum = np.cumsum(hist)
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, nranks):
        target = total * r / nranks
        c = int(np.searchsorted(cum, target, side="left")) + 1          # first cut with cum[c-1] >= target ...
        if c >= 2 and abs(cum[c - 2] - target) < abs(cum[c - 1] - target):
            c -= 1                                                        # ... or the plane before, whichever is nearer
        c = max(c, cuts[-1] + minwidth)
        c = min(c, ncells_x - (nranks - r) * minwidth)
        cuts.append(c)
    cuts.append(ncells_x)
    return cuts


class LocalTransport:
    """All ranks live in this process (same device); exchanges are plain device copies.  For tests."""

    def __init__(self, nranks):
        self.nranks = nranks
        self.mail = {}

    def exchange_counts(self, rank, nleft, nright):
        self.mail[("cnt", rank)] = (nleft, nright)

    def counts_from(self, rank):
        left = self.mail[("cnt", rank - 1)][1] if rank > 0 else 0
        right = self.mail[("cnt", rank + 1)][0] if rank < self.nranks - 1 else 0
        return left, right


class SlabRank:
    """One slab: a Simulation restricted to [cxa,cxb) + its exchange buffers (torch uint8 tensors on the device)."""

    def __init__(self, casepath, options, rank, nranks, device=0):
        import torch
        self.torch = torch
        self.rank, self.nranks = rank, nranks
        self.dev = torch.device("cuda", device)
        self.sim = L.Simulation(casepath, options=options, device=device, slab=(rank, nranks))   # uploads only this slab
        self.ctx = self.sim.ctx
        self.p = self.sim.params
        self.cuts, maxplane = self.sim.slab_info(nranks)
        self.cxa, self.cxb = self.cuts[rank], self.cuts[rank + 1]
        # exchange buffers: scelldiv cell planes of halo per side, or the migrants of one step (far fewer)
        self.halo_planes = int(self.p.scelldiv) * (2 if self.p.tboundary == L.BC_MDBC else 1)   # csrc/slab.cu slab_halo_planes
        self.cap_records = int(2 * self.halo_planes * maxplane + 4096)
        self.symplectic = (self.p.tstep == L.STEP_SYMPLECTIC)
        maxbytes = 48 + 48 + 16 + 16 + 32     # largest record csrc/slab.cu can produce (the size varies with the step phase)
        self._bufbytes = self.cap_records * maxbytes
        self._bufs = None        # allocated on first use: the native step keeps its own buffers inside the library
        self._ext = None
        self._dt_t = None
        self._ft_t = None

    def _buf(self, k):
        if self._bufs is None:
            self._bufs = [self.torch.empty(self._bufbytes, dtype=self.torch.uint8, device=self.dev) for _ in range(4)]
        return self._bufs[k]

    send_l = property(lambda self: self._buf(0))
    send_r = property(lambda self: self._buf(1))
    recv_l = property(lambda self: self._buf(2))
    recv_r = property(lambda self: self._buf(3))

    def on_stream(self):
        """Context that makes the ctx's CUDA stream torch's current stream: NCCL operations issued inside are ordered with
        the ctx's kernels by stream semantics (no host synchronisation between pack -> send/recv -> unpack)."""
        if self._ext is None:
            self._ext = self.torch.cuda.ExternalStream(self.ctx.stream(), device=self.dev)
        return self.torch.cuda.stream(self._ext)

    def pack(self, kind):
        return self.ctx.slab_pack(kind, self.send_l.data_ptr(), self.send_r.data_ptr(), self.cap_records)

    def pack_both(self):
        return self.ctx.slab_pack_both(self.send_l.data_ptr(), self.send_r.data_ptr(), self.send_l.numel())

    def unpack_both(self, buf, nmig, nhalo):
        self.ctx.slab_unpack_both(buf.data_ptr(), nmig, nhalo)

    def unpack(self, kind, nl, nr):
        if nl:
            self.ctx.slab_unpack(kind, self.recv_l.data_ptr(), nl)
        if nr:
            self.ctx.slab_unpack(kind, self.recv_r.data_ptr(), nr)

    def sync(self):
        self.torch.cuda.synchronize(self.dev)

    def dt_tensor(self):
        """The three dt maxima as an int32 view over the ctx's device memory (zero-copy, via __cuda_array_interface__)."""
        if self._dt_t is None:
            ptr = self.ctx.slab_dt_buffer()

            class _W:
                __cuda_array_interface__ = {"shape": (3,), "typestr": "<i4", "data": (ptr, False), "version": 2}
            self._dt_t = self.torch.as_tensor(_W(), device=self.dev)
        return self._dt_t

    def close(self):
        self.sim.close()


def _exchange_local(ranks, kind):
    """Pack on every rank, copy buffers to the neighbours, unpack (single-process transport)."""
    counts = [r.pack(kind) for r in ranks]
    for r in ranks:
        r.sync()
    rb = ranks[0].ctx.slab_record_bytes(kind)
    for i, r in enumerate(ranks):
        nl = counts[i - 1][1] if i > 0 else 0
        nr = counts[i + 1][0] if i < len(ranks) - 1 else 0
        if nl:
            r.recv_l[:nl * rb].copy_(ranks[i - 1].send_r[:nl * rb])
        if nr:
            r.recv_r[:nr * rb].copy_(ranks[i + 1].send_l[:nr * rb])
        r.sync()
        r.unpack(kind, nl, nr)
    return counts


def _exchange_dist(rank, kind, dist):
    """Pack, exchange counts and payloads with the two neighbours (batched isend/irecv), unpack.  Everything is issued on
    the ctx's stream (``rank.on_stream()``), so the only host synchronisations are the two that make counts visible to
    the host: the packed counts (inside b2s_slab_pack) and the neighbours' counts."""
    torch = rank.torch
    me, n = rank.rank, rank.nranks
    with rank.on_stream():
        nl_s, nr_s = rank.pack(kind)
        rb = rank.ctx.slab_record_bytes(kind)
        cnt_send = torch.tensor([nl_s, nr_s], dtype=torch.int64, device=rank.dev)
        cnt_recv = torch.zeros(2, dtype=torch.int64, device=rank.dev)
        ops = []
        if me > 0:
            ops += [dist.P2POp(dist.isend, cnt_send[0:1], me - 1), dist.P2POp(dist.irecv, cnt_recv[0:1], me - 1)]
        if me < n - 1:
            ops += [dist.P2POp(dist.isend, cnt_send[1:2], me + 1), dist.P2POp(dist.irecv, cnt_recv[1:2], me + 1)]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        nl_r, nr_r = (int(v) for v in cnt_recv.tolist())
        ops = []
        if me > 0:
            if nl_s:
                ops.append(dist.P2POp(dist.isend, rank.send_l[:nl_s * rb], me - 1))
            if nl_r:
                ops.append(dist.P2POp(dist.irecv, rank.recv_l[:nl_r * rb], me - 1))
        if me < n - 1:
            if nr_s:
                ops.append(dist.P2POp(dist.isend, rank.send_r[:nr_s * rb], me + 1))
            if nr_r:
                ops.append(dist.P2POp(dist.irecv, rank.recv_r[:nr_r * rb], me + 1))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        rank.unpack(kind, nl_r, nr_r)
    return (nl_s, nr_s, nl_r, nr_r)


def _exchange_both_local(ranks):
    """Merged migrant+halo exchange, single-process transport (device copies)."""
    counts = [r.ctx.slab_pack_both(r.send_l.data_ptr(), r.send_r.data_ptr(), r.send_l.numel()) for r in ranks]
    for r in ranks:
        r.sync()
    rbm, rbh = ranks[0].ctx.slab_record_bytes(MIGRANT), ranks[0].ctx.slab_record_bytes(HALO)
    for i, r in enumerate(ranks):
        if i > 0:                                   # what the left neighbour sent to its right
            nm, nh = counts[i - 1][1], counts[i - 1][3]
            nb = nm * rbm + nh * rbh
            if nb:
                r.recv_l[:nb].copy_(ranks[i - 1].send_r[:nb]); r.sync()
                r.ctx.slab_unpack_both(r.recv_l.data_ptr(), nm, nh)
        if i < len(ranks) - 1:
            nm, nh = counts[i + 1][0], counts[i + 1][2]
            nb = nm * rbm + nh * rbh
            if nb:
                r.recv_r[:nb].copy_(ranks[i + 1].send_l[:nb]); r.sync()
                r.ctx.slab_unpack_both(r.recv_r.data_ptr(), nm, nh)
    return counts, rbm, rbh


def _exchange_both_dist(rank, dist):
    """Merged migrant+halo exchange over torch.distributed: ONE count exchange and ONE payload exchange per interaction,
    everything on the ctx's stream; the host waits twice (packed counts, neighbours' counts)."""
    torch = rank.torch
    me, n = rank.rank, rank.nranks
    with rank.on_stream():
        c = rank.pack_both()
        rbm, rbh = rank.ctx.slab_record_bytes(MIGRANT), rank.ctx.slab_record_bytes(HALO)
        cnt_send = torch.tensor([c[0], c[2], c[1], c[3]], dtype=torch.int64, device=rank.dev)   # (mig,halo) left, (mig,halo) right
        cnt_recv = torch.zeros(4, dtype=torch.int64, device=rank.dev)
        ops = []
        if me > 0:
            ops += [dist.P2POp(dist.isend, cnt_send[0:2], me - 1), dist.P2POp(dist.irecv, cnt_recv[0:2], me - 1)]
        if me < n - 1:
            ops += [dist.P2POp(dist.isend, cnt_send[2:4], me + 1), dist.P2POp(dist.irecv, cnt_recv[2:4], me + 1)]
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        ml, hl, mr, hr = (int(v) for v in cnt_recv.tolist())
        sl, sr = c[0] * rbm + c[2] * rbh, c[1] * rbm + c[3] * rbh
        rl, rr = ml * rbm + hl * rbh, mr * rbm + hr * rbh
        ops = []
        if me > 0:
            if sl:
                ops.append(dist.P2POp(dist.isend, rank.send_l[:sl], me - 1))
            if rl:
                ops.append(dist.P2POp(dist.irecv, rank.recv_l[:rl], me - 1))
        if me < n - 1:
            if sr:
                ops.append(dist.P2POp(dist.isend, rank.send_r[:sr], me + 1))
            if rr:
                ops.append(dist.P2POp(dist.irecv, rank.recv_r[:rr], me + 1))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        if rl:
            rank.unpack_both(rank.recv_l, ml, hl)
        if rr:
            rank.unpack_both(rank.recv_r, mr, hr)
    return c, rbm, rbh


class SlabSimulation:
    """Drives N slabs through ComputeStep_Sym / ComputeStep_Ver (JSphGpuSingle.cpp:761-819) with the exchanges in between.

    ``dist`` = torch.distributed (already initialised, NCCL): this process owns ONE rank.
    ``dist`` = None: all ranks are created in this process on ``device`` (LocalTransport, for tests)."""

    def __init__(self, casepath, options="", nranks=1, dist=None, device=0, recut_every=0, merged=True, native=None):
        """``native``: the whole step runs inside the library (csrc/slab_comm.cu: ncclSend/ncclRecv/ncclAllReduce issued from
        C++, counts kept on the device, one host synchronisation per step); Python only hands over the NCCL id.  Default: on
        whenever the slabs are wider than the halo (what the device-counted exchange needs).  ``native=False`` keeps the
        round-1 orchestration in this file (torch.distributed P2P ops, two host waits per exchange); with ``dist=None`` the
        native path uses the library's local test transport."""
        import torch
        self.torch, self.dist = torch, dist
        if dist is None:
            self.ranks = [SlabRank(casepath, options, r, nranks, device=device) for r in range(nranks)]
        else:
            self.ranks = [SlabRank(casepath, options, dist.get_rank(), nranks, device=device)]
        self.symplectic = self.ranks[0].symplectic
        self.exchanged_bytes = 0
        self.migrated = 0
        self.merged = bool(merged)              # migrants and halos in one message per neighbour (csrc/slab.cu pack_both)
        self.recut_every = int(recut_every)     # >0: rebalance the cuts every that many steps (dynamic load balance)
        self.nstep = 0
        import os
        self._prof = [] if (os.environ.get("B2S_SLAB_PROFILE") and dist is not None) else None
        self.cuts = self.ranks[0].cuts
        if native is None:
            native = os.environ.get("B2S_SLAB_NATIVE", "1") != "0"
        self.native = bool(native) and nranks > 1 and self._merged_ok()
        if self.native:
            _, maxplane = self.ranks[0].sim.slab_info(nranks)
            if dist is None:
                L.slab_comm_init_local([r.ctx for r in self.ranks], maxplane)
            else:
                # rank 0 makes the NCCL id; it travels once through torch.distributed (plumbing), the data plane is the library's
                ids = [L.slab_comm_id() if dist.get_rank() == 0 else None]
                dist.broadcast_object_list(ids, src=0)
                self.ranks[0].ctx.slab_comm_init(ids[0], maxplane)
            if self.recut_every:
                for r in self.ranks:
                    r.ctx.slab_comm_set_recut(self.recut_every)
            self.ranks[0].ctx.slab_exchange()
            return
        self._exchange(first=True)

    # migrants then halos, then the cell divide (RunCellDivide with the ghosts in place)
    def _merged_ok(self):
        """The merged exchange needs every slab wider than the halo (cuts are the same on all ranks)."""
        hw = self.ranks[0].halo_planes
        return self.merged and all(b - a > hw for a, b in zip(self.cuts, self.cuts[1:]))

    def _exchange(self, first=False):
        if not first and self._merged_ok():
            if self.dist is None:
                cs, rbm, rbh = _exchange_both_local(self.ranks)
            else:
                c, rbm, rbh = _exchange_both_dist(self.ranks[0], self.dist)
                cs = [c]
            for c in cs:
                self.exchanged_bytes += (c[0] + c[1]) * rbm + (c[2] + c[3]) * rbh
                self.migrated += c[0] + c[1]
            for r in self.ranks:
                r.ctx.celldiv()
            return
        for kind in ((HALO,) if first else (MIGRANT, HALO)):
            if self.dist is None:
                c = _exchange_local(self.ranks, kind)
                rb = self.ranks[0].ctx.slab_record_bytes(kind)
                self.exchanged_bytes += sum(a + b for a, b in c) * rb
                if kind == MIGRANT:
                    self.migrated += sum(a + b for a, b in c)
            else:
                c = _exchange_dist(self.ranks[0], kind, self.dist)
                self.exchanged_bytes += (c[0] + c[1]) * self.ranks[0].ctx.slab_record_bytes(kind)
                if kind == MIGRANT:
                    self.migrated += c[0] + c[1]
        for r in self.ranks:
            r.ctx.celldiv()

    def _allreduce_dt(self):
        if self.dist is None:
            ts = [r.dt_tensor() for r in self.ranks]
            for r in self.ranks:
                r.sync()
            m = self.torch.stack(ts).max(dim=0).values
            for t in ts:
                t.copy_(m)
            for r in self.ranks:
                r.sync()
        else:
            r = self.ranks[0]
            with r.on_stream():     # ordered after the interaction kernels and before k_dtvariable by the stream itself
                self.dist.all_reduce(r.dt_tensor(), op=self.dist.ReduceOp.MAX)

    def _floating(self, predictor):
        """RunFloating across slabs: local sums -> all-reduce(SUM) -> rigid-body update on every rank."""
        torch = self.torch
        views = []
        for r in self.ranks:
            ptr, n = r.ctx.slab_ft_sums()
            if not n:
                return
            if r._ft_t is None:
                class _W:
                    __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
                r._ft_t = torch.as_tensor(_W(), device=r.dev)
            views.append(r._ft_t)
        if self.dist is None:
            for r in self.ranks:
                r.sync()
            tot = torch.stack(views).sum(dim=0)
            for v in views:
                v.copy_(tot)
            for r in self.ranks:
                r.sync()
        else:
            with self.ranks[0].on_stream():
                self.dist.all_reduce(views[0], op=self.dist.ReduceOp.SUM)
        for r in self.ranks:
            r.ctx.slab_ft_apply(predictor)

    def _interaction(self, interstep):
        for r in self.ranks:
            r.ctx.slab_forces(interstep)
        self._allreduce_dt()
        for r in self.ranks:
            r.ctx.slab_apply_dt(interstep)

    def _mark(self, name):
        """B2S_SLAB_PROFILE=1: CUDA events on the ctx stream at the phase boundaries of a step (rank 0 of this process)."""
        if self._prof is None:
            return
        ev = self.torch.cuda.Event(enable_timing=True)
        with self.ranks[0].on_stream():
            ev.record()
        self._prof.append((name, ev))

    def profile_report(self):
        """Mean milliseconds per phase since construction (None unless B2S_SLAB_PROFILE=1)."""
        if self.native:
            txt = self.ranks[0].ctx.slab_comm_profile()
            return txt or None
        if not self._prof:
            return None
        self.torch.cuda.synchronize()
        acc, cnt = {}, {}
        for (n0, e0), (n1, e1) in zip(self._prof, self._prof[1:]):
            if n1 == "begin":
                continue
            acc[n1] = acc.get(n1, 0.0) + e0.elapsed_time(e1)
            cnt[n1] = cnt.get(n1, 0) + 1
        nsteps = max(cnt.get("step_end", 1), 1)
        return {k: v / nsteps for k, v in acc.items()}

    def comm_stats(self, reset=False):
        """Native path: (ms inside pack+transfer+unpack on rank 0 of this process, exchanges, bytes sent, re-cuts)."""
        return self.ranks[0].ctx.slab_comm_stats(reset) if self.native else (0.0, 0, 0, 0, 0)

    def migrated_total(self) -> int:
        """Particles that changed owner so far (this process's ranks)."""
        if self.native:
            return sum(r.ctx.slab_comm_stats()[4] for r in self.ranks)
        return self.migrated

    def run_steps(self, n: int):
        if self.native:
            self.ranks[0].ctx.slab_run_steps(n)
            self.nstep += n
        else:
            for _ in range(n):
                self.step()

    def step(self, final_exchange=True) -> float:
        if self.native and final_exchange:
            dt = self.ranks[0].ctx.slab_step()      # local transport: one call drives every rank of the group
            self.nstep += 1
            if self.recut_every and self.nstep % self.recut_every == 0:
                self._refresh_cuts()
            return dt
        self._mark("begin")
        if self.symplectic:
            self._interaction(L.INTER_SYMPREDICTOR)
            self._mark("interaction_pre")
            for r in self.ranks:
                r.ctx.slab_integrate(1)
            self._floating(True)
            self._mark("integrate_pre")
            self._exchange()
            self._mark("exchange_mid+celldiv")
            self._interaction(L.INTER_SYMCORRECTOR)
            self._mark("interaction_corr")
            for r in self.ranks:
                r.ctx.slab_integrate(2)
            self._floating(False)
            self._mark("integrate_corr")
        else:
            self._interaction(L.INTER_VERLET)
            self._mark("interaction")
            for r in self.ranks:
                r.ctx.slab_integrate(0)
            self._floating(False)
            self._mark("integrate")
        dts = [r.ctx.slab_step_end() for r in self.ranks]
        self.nstep += 1
        self._mark("step_end")
        if final_exchange:
            if self.recut_every and self.nstep % self.recut_every == 0:
                self._exchange()          # settle this step's migrants against the old cuts first
                self.recut()
            else:
                self._exchange()
            self._mark("exchange_end+celldiv")
        return dts[0]

    def _refresh_cuts(self):
        for r in self.ranks:
            r.cxa, r.cxb = r.ctx.slab_get_range()
        if self.dist is None:
            self.cuts = [self.ranks[0].cxa] + [r.cxb for r in self.ranks]
            for r in self.ranks:
                r.cuts = self.cuts

    def recut(self, cuts=None):
        """Dynamic load balance: new cut planes from the global particle histogram along x (all-reduced), then the particles
        outside their slab's new range migrate (one neighbour hop per round, repeated until nobody has migrants left), then
        halos + divide.  Call between steps.  ``cuts`` overrides the balanced cuts (tests)."""
        if self.native and cuts is None:
            self.ranks[0].ctx.slab_recut()
            self._refresh_cuts()
            return self.cuts
        torch = self.torch
        ncx = int(self.ranks[0].p.domcells[0])
        nranks = self.ranks[0].nranks
        for r in self.ranks:                      # the halos are rebuilt at the end: the ghosts of the last exchange go
            r.ctx.slab_drop_ghosts()
            r.ctx.celldiv()
        if cuts is None:
            hists = []
            for r in self.ranks:
                h = torch.zeros(ncx, dtype=torch.int32, device=r.dev)
                with r.on_stream():
                    r.ctx.slab_histogram(h.data_ptr())
                    if self.dist is not None:
                        self.dist.all_reduce(h, op=self.dist.ReduceOp.SUM)
                hists.append(h)
            for r in self.ranks:
                r.sync()
            total = hists[0].cpu().numpy().astype(np.int64) if self.dist is not None else sum(h.cpu().numpy().astype(np.int64) for h in hists)
            cuts = compute_cuts_hist(total, nranks, self.ranks[0].halo_planes)
        self.cuts = [int(c) for c in cuts]
        for r in self.ranks:
            r.cuts = self.cuts
            r.cxa, r.cxb = self.cuts[r.rank], self.cuts[r.rank + 1]
            r.ctx.slab_set_range(r.cxa, r.cxb)
        for _ in range(nranks):                    # a particle travels at most nranks-1 hops
            if self.dist is None:
                c = _exchange_local(self.ranks, MIGRANT)
                moved = sum(a + b for a, b in c)
            else:
                c = _exchange_dist(self.ranks[0], MIGRANT, self.dist)
                t = torch.tensor([c[0] + c[1]], dtype=torch.int64, device=self.ranks[0].dev)
                self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
                moved = int(t.item())
            self.migrated += moved if self.dist is None else c[0] + c[1]
            if moved == 0:
                break
            for r in self.ranks:                   # forwarded particles must be (re)flagged against the new range
                r.ctx.celldiv()
        else:
            raise RuntimeError("recut: migrants left after nranks rounds")
        self._exchange(first=True)
        if self.native and not self._merged_ok():
            self.native = False          # forced cuts no wider than the halo: back to the host-counted exchange of this file
        return self.cuts

    def host_roundtrip_begin(self):
        """After ``step(final_exchange=False)``: drop the ghosts (they were marked by the integrator) so that a download
        returns exactly the particles this rank owns (would-be migrants included)."""
        for r in self.ranks:
            r.ctx.celldiv()

    def host_roundtrip_end(self):
        """After re-uploading the owned particles: move migrants, rebuild halos, divide."""
        if self.native:
            self.ranks[0].ctx.slab_exchange()
        else:
            self._exchange()

    def set_dirout(self, dirout: str):
        """PART output of a multi-GPU run: every rank writes its own piece (Part_pRR_XXXX.bi4, the reference's piece format,
        JPartDataBi4.cpp:99-127); nothing is gathered."""
        for r in self.ranks:
            r.sim.set_dirout(dirout)

    def save_part(self):
        """Writes the next PART: a rank saves the particles it owns (halo ghosts carry CODE_PERIODIC and are skipped by the
        download), so every particle of the case lands in exactly one piece."""
        for r in self.ranks:
            r.sim.save_part()

    def gather_byid(self):
        """Owned particles of the local ranks, sorted by id: (idp, pos, velrho)."""
        ids, pos, vr = [], [], []
        for r in self.ranks:
            code = r.ctx.array("code")
            own = (code & 0xe000) == 0
            ids.append(r.ctx.array("idp")[own]); pos.append(r.ctx.array("pos")[own]); vr.append(r.ctx.array("velrho")[own])
        ids, pos, vr = np.concatenate(ids), np.concatenate(pos), np.concatenate(vr)
        o = np.argsort(ids, kind="stable")
        return ids[o], pos[o], vr[o]

    def owned_counts(self):
        out = []
        for r in self.ranks:
            code = r.ctx.array("code")
            out.append(int(((code & 0xe000) == 0).sum()))
        return out

    def close(self):
        for r in self.ranks:
            r.close()


def cuts_from_ctx(ctx, params, nranks):
    """Balanced cuts from the particles a freshly loaded context holds (every rank loads the same case, so every rank
    derives the same cuts without communication).  Cell x comes from dcell (JDsDcellDef.h:38 DCEL_Cellx)."""
    dcell = ctx.array("dcell")
    dcc = int(params.domcellcode)
    cellx = dcell[dcell != 0xFFFFFFFF] >> ((dcc >> 10) & 31)
    return compute_cuts(cellx, int(params.domcells[0]), nranks)
