"""Multi-GPU slab run over the real transport (NCCL, one process per GPU) against the single-context run.
Needs at least 2 GPUs: skipped on the single-GPU test box, run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu`."""
import os
import socket
import subprocess
import sys

import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("options,nsteps,recut", [("-symplectic -cellmode:half", 12, ""), ("-verlet", 12, ""),
                                                    ("-symplectic -cellmode:half", 400, "recut")])
def test_nccl_slabs_match_single_context(options, nsteps, recut, tmp_path, built_lib):
    H.write_case(tmp_path, dp=0.04 if recut else 0.03)
    world = min(_ngpu(), 4) if recut else 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    worker = os.path.join(os.path.dirname(__file__), "slab_nccl_worker.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), worker, str(tmp_path / "Case"), options, str(nsteps)] + ([recut] if recut else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    print(r.stdout[-1500:])
    assert r.returncode == 0 and "SLAB_NCCL_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-4000:]
