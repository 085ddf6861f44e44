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


def _wait_all(procs, limit=600):
    """Output of every rank; a rank that exits with an error takes the others down (they would wait in a collective for ever)."""
    import time
    t0 = time.time()
    while any(p.poll() is None for p in procs):
        if any(p.poll() not in (None, 0) for p in procs) or time.time() - t0 > limit:
            for p in procs:
                if p.poll() is None:
                    p.kill()
            break
        time.sleep(0.2)
    return [p.communicate()[0] for p in procs]


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


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_cli_two_ranks_write_pieces(tmp_path, built_lib):
    """The C++ command line on two GPUs: one b200sph process per GPU (RANK / WORLD_SIZE / LOCAL_RANK in the environment, the NCCL id
    through a file), collective steps inside the library (csrc/slab_comm.cu), one piece of every PART per rank.  The union of the
    pieces must be the particle set of the single-GPU run of the same command, positions to rounding."""
    import numpy as np
    from dualsphysics_b200 import bi4
    H.write_case(tmp_path, dp=0.03, params={"TimeMax": 0.02, "TimeOut": 0.005})
    exe = os.path.join(H.ROOT, "dualsphysics_b200", "lib", "b200sph")
    opts = ["-symplectic", "-cellmode:half", "-tmax:0.01", "-tout:0.005"]
    one = subprocess.run([exe, str(tmp_path / "Case"), str(tmp_path / "out1")] + opts, capture_output=True, text=True, timeout=600)
    assert one.returncode == 0, one.stdout[-2000:] + one.stderr[-2000:]
    idfile = str(tmp_path / "nccl.id")
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), B2S_NCCL_ID_FILE=idfile)
        procs.append(subprocess.Popen([exe, str(tmp_path / "Case"), str(tmp_path / "out2")] + opts, env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = _wait_all(procs)
    assert all(p.returncode == 0 for p in procs), "\n".join(o[-2000:] for o in outs)
    assert "Finished execution (code=0)" in outs[0]
    k = 0
    while os.path.exists(str(tmp_path / "out1" / "data" / f"Part_{k:04d}.bi4")):
        ref = bi4.load_part(str(tmp_path / "out1" / "data" / f"Part_{k:04d}.bi4"))
        pieces = [bi4.load_part(str(tmp_path / "out2" / "data" / f"Part_p{r:02d}_{k:04d}.bi4")) for r in range(2)]
        ids = np.concatenate([p["Idp"] for p in pieces]); pos = np.concatenate([p["pos"] for p in pieces])
        o, ro = np.argsort(ids, kind="stable"), np.argsort(ref["Idp"], kind="stable")
        assert np.array_equal(ids[o], ref["Idp"][ro]), f"PART {k}: the pieces do not add up to the particle set"
        assert float(np.max(np.abs(pos[o] - ref["pos"][ro]))) < 1e-6, k
        k += 1
    assert k >= 3, k


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_cli_two_ranks_gauges(tmp_path, built_lib):
    """Gauges in a slab-decomposed run of the command line (csrc/slab_comm.cu grp_gauges over NCCL; rank 0 writes the CSV files): the
    reference's gauge fixture (a velocity point, a vertical SWL line, an SWL line along x that crosses the cut) on two GPUs against
    the single-GPU run of the same command: the same rows and times, values within 1e-4 of their column's scale (the two runs differ by
    2e-6 of max|v| in the state, tests/test_gpu_slab.py, and the files print six digits)."""
    import json
    import numpy as np
    fx = json.load(open(os.path.join(H.GOLDEN, "gauges_ref", "fixture_gauges.json")))
    H.write_case(tmp_path, **fx["spec"])
    exe = os.path.join(H.ROOT, "dualsphysics_b200", "lib", "b200sph")
    opts = list(fx["options"]) + ["-cellmode:half", "-tmax:0.02"]
    one = subprocess.run([exe, str(tmp_path / "Case"), str(tmp_path / "out1")] + opts, capture_output=True, text=True, timeout=600)
    assert one.returncode == 0, one.stdout[-2000:] + one.stderr[-2000:]
    idfile = str(tmp_path / "nccl.id")
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), B2S_NCCL_ID_FILE=idfile)
        procs.append(subprocess.Popen([exe, str(tmp_path / "Case"), str(tmp_path / "out2")] + opts, env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = _wait_all(procs)
    assert all(p.returncode == 0 for p in procs), "\n".join(o[-2000:] for o in outs)
    names = sorted(f for f in os.listdir(tmp_path / "out1") if f.startswith("Gauges") and f.endswith(".csv"))
    assert len(names) == 3, names
    for f in names:
        a, b = open(tmp_path / "out1" / f).read().strip().splitlines(), open(tmp_path / "out2" / f).read().strip().splitlines()
        assert a[0] == b[0] and len(a) == len(b) and len(a) > 5, (f, len(a), len(b))
        x = np.array([[float(v) for v in row.split(";")] for row in a[1:]]); y = np.array([[float(v) for v in row.split(";")] for row in b[1:]])
        assert np.array_equal(x[:, 0], y[:, 0]), f                       # the same times
        scale = np.maximum(np.abs(x).max(axis=0), 1e-3)
        assert np.max(np.abs(x - y) / scale) <= 1e-4, (f, float(np.max(np.abs(x - y) / scale)))


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_cli_two_ranks_partout(tmp_path, built_lib):
    """Excluded particles of a slab-decomposed run: every rank writes what its own particles lose as PartOut_pRR_000.obi4 (the
    reference's piece naming, JPartOutBi4Save.cpp:102-113).  A lower density limit just under rho0 (-rhopout:999.5:1300, the window of the reference's
    list-file fixture) excludes particles early in the dam break; the pieces of the two ranks together hold the particles the single-GPU run excluded (a particle whose density sits at the
    limit to rounding may differ: at most 2% of them), and PART pieces + excluded particles still add up to the case."""
    import numpy as np
    from dualsphysics_b200 import bi4
    H.write_case(tmp_path, dp=0.03, params={"TimeMax": 0.02, "TimeOut": 0.005})
    exe = os.path.join(H.ROOT, "dualsphysics_b200", "lib", "b200sph")
    opts = ["-symplectic", "-cellmode:half", "-tmax:0.015", "-tout:0.005", "-rhopout:999.5:1300"]
    one = subprocess.run([exe, str(tmp_path / "Case"), str(tmp_path / "out1")] + opts, capture_output=True, text=True, timeout=600)
    assert one.returncode == 0, one.stdout[-2000:] + one.stderr[-2000:]
    idfile = str(tmp_path / "nccl.id")
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), B2S_NCCL_ID_FILE=idfile)
        procs.append(subprocess.Popen([exe, str(tmp_path / "Case"), str(tmp_path / "out2")] + opts, env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = _wait_all(procs)
    assert all(p.returncode == 0 for p in procs), "\n".join(o[-2000:] for o in outs)

    def out_ids(path):
        items = bi4.load_list(path)
        return items[0], np.concatenate([np.asarray(it.array("Idp")) for it in items[1:]] + [np.zeros(0, np.uint32)])
    h1, ids1 = out_ids(str(tmp_path / "out1" / "data" / "PartOut_000.obi4"))
    pieces = [out_ids(str(tmp_path / "out2" / "data" / f"PartOut_p{r:02d}_000.obi4")) for r in range(2)]
    assert [h.get("Piece") for h, _ in pieces] == [0, 1] and all(h.get("Npiece") == 2 for h, _ in pieces)
    assert [(n, t) for n, t, _ in pieces[0][0].values] == [(n, t) for n, t, _ in h1.values]
    ids2 = np.concatenate([i for _, i in pieces])
    assert len(ids1) > 5, len(ids1)                                    # the window does exclude particles
    assert len(np.unique(ids2)) == len(ids2)                           # nobody is excluded twice
    assert len(np.setxor1d(ids1, ids2)) <= max(2, len(ids1) // 50), (len(ids1), len(ids2), len(np.setxor1d(ids1, ids2)))
    # last PART: the pieces and everything excluded so far partition the case
    k = 0
    while os.path.exists(str(tmp_path / "out2" / "data" / f"Part_p00_{k + 1:04d}.bi4")):
        k += 1
    last = np.concatenate([bi4.load_part(str(tmp_path / "out2" / "data" / f"Part_p{r:02d}_{k:04d}.bi4"))["Idp"] for r in range(2)])
    allids = np.concatenate([last, ids2])
    assert len(np.unique(allids)) == len(allids) == len(bi4.load_part(str(tmp_path / "Case.bi4"))["Idp"])
