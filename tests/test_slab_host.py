"""CPU tests of the multi-GPU host logic (no GPU, gloo backend): cut computation and the neighbour exchange protocol."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import helpers as H  # noqa: F401
from dualsphysics_b200 import slab


def test_compute_cuts_balances_particles():
    rng = np.random.default_rng(3)
    cx = np.floor(rng.random(200000) ** 3 * 64)
    for n in (1, 2, 4, 8):
        cuts = slab.compute_cuts(cx, 64, n)
        assert cuts[0] == 0 and cuts[-1] == 64 and len(cuts) == n + 1
        assert all(b > a for a, b in zip(cuts, cuts[1:]))
        own = np.array([((cx >= a) & (cx < b)).sum() for a, b in zip(cuts, cuts[1:])])
        assert own.sum() == len(cx)
        if n > 1:   # a whole plane is the granularity: the densest plane bounds the imbalance
            assert own.max() <= len(cx) / n + np.bincount(cx.astype(int)).max()


def test_compute_cuts_min_width():
    rng = np.random.default_rng(5)
    cx = np.floor(rng.random(5000) ** 4 * 40)
    cuts = slab.compute_cuts(cx, 40, 6, minwidth=2)
    assert cuts[0] == 0 and cuts[-1] == 40 and all(b - a >= 2 for a, b in zip(cuts, cuts[1:]))


def test_compute_cuts_degenerate():
    # fewer populated planes than ranks: every slab still gets at least one plane
    assert slab.compute_cuts(np.zeros(10), 4, 4) == [0, 1, 2, 3, 4]
    assert slab.compute_cuts(np.full(10, 7), 8, 3)[-1] == 8


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_protocol_gloo(world):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    worker = os.path.join(os.path.dirname(__file__), "slab_gloo_worker.py")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), worker],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "SLAB_GLOO_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
