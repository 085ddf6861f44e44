#!/usr/bin/env python
"""GPU checks of the host features added after the last GPU minute of round 1 (run by tools/gpu_r2_first.sh; they become `-m gpu` tests
once they have passed on a B200):
  1. mDBC restart: a run restarted from PART 1 through PartExtra_0001.bi4 against the uninterrupted run (like test_restart_from_part);
  2. the multi-GPU runner on one GPU (`python -m dualsphysics_b200.run`): same PART files as the C++ command line;
  3. piece output of two slab ranks living on one GPU: the pieces of a PART hold every particle once, positions equal to a single-GPU run.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H  # noqa: E402
from dualsphysics_b200 import bi4, slab  # noqa: E402

CLI = os.path.join(ROOT, "dualsphysics_b200", "lib", "b200sph")


def sh(*cmd):
    r = subprocess.run(list(cmd), capture_output=True, text=True)
    if r.returncode:
        print(r.stdout[-1500:], r.stderr[-1500:])
    return r.returncode


def check_mdbc_restart(td):
    d = os.path.join(td, "mdbc"); os.makedirs(d)
    H.write_case(d, dp=0.04, layers=4, normals=True, params={"Boundary": 2, "SlipMode": 1, "TimeMax": 0.03, "TimeOut": 0.01})
    case = os.path.join(d, "Case")
    assert sh(CLI, case, os.path.join(d, "A"), "-symplectic", "-svextraparts:1") == 0
    assert sh(CLI, case, os.path.join(d, "B"), "-symplectic", "-partbegin:1", os.path.join(d, "A", "data")) == 0
    worst = 0.0
    for k in (1, 2, 3):
        a, b = bi4.load_part(os.path.join(d, "A", "data", f"Part_{k:04d}.bi4")), bi4.load_part(os.path.join(d, "B", "data", f"Part_{k:04d}.bi4"))
        assert np.array_equal(a["Idp"], b["Idp"])
        worst = max(worst, float(np.max(np.abs(a["pos"] - b["pos"]))))
    print(f"[1] mDBC restart: max |dpos| vs uninterrupted run = {worst:.3e} m")
    return worst < 1e-7


def check_runner_one_gpu(td):
    d = os.path.join(td, "runner"); os.makedirs(d)
    H.write_case(d, dp=0.04, params={"TimeMax": 0.02, "TimeOut": 0.01})
    case = os.path.join(d, "Case")
    assert sh(CLI, case, os.path.join(d, "A"), "-symplectic") == 0
    assert sh(sys.executable, "-m", "dualsphysics_b200.run", case, os.path.join(d, "B"), "-symplectic") == 0
    worst = 0.0
    for k in (0, 1, 2):
        a, b = bi4.load_part(os.path.join(d, "A", "data", f"Part_{k:04d}.bi4")), bi4.load_part(os.path.join(d, "B", "data", f"Part_{k:04d}.bi4"))
        assert np.array_equal(a["Idp"], b["Idp"]) and abs(a["part"]["TimeStep"] - b["part"]["TimeStep"]) < 1e-9
        worst = max(worst, float(np.max(np.abs(a["pos"] - b["pos"]))))
    print(f"[2] runner on one GPU: max |dpos| vs command line = {worst:.3e} m")
    return worst < 1e-7


def check_pieces_two_ranks(td):
    d = os.path.join(td, "pieces"); os.makedirs(d)
    H.write_case(d, dp=0.03, params={"TimeMax": 0.02, "TimeOut": 0.01})
    case = os.path.join(d, "Case")
    assert sh(CLI, case, os.path.join(d, "A"), "-symplectic", "-nsteps:10", "-svsteps:1") == 0
    ss = slab.SlabSimulation(case, options="-symplectic", nranks=2, dist=None, device=0)
    ss.set_dirout(os.path.join(d, "B"))
    ss.save_part()
    for _ in range(10):
        ss.step()
    ss.save_part()
    ss.close()
    ref = bi4.load_part(os.path.join(d, "A", "data", "Part_0010.bi4"))
    ids, pos = [], []
    for r in range(2):
        root = bi4.load(os.path.join(d, "B", "data", f"Part_p{r:02d}_0001.bi4"))
        part = root.item("PART_0001")
        ids.append(np.asarray(part.array("Idp"))); pos.append(np.asarray(part.array("Pos" if part.array("Posd") is None else "Posd")).reshape(-1, 3))
    ids, pos = np.concatenate(ids), np.concatenate(pos)
    o = np.argsort(ids)
    assert np.array_equal(ids[o], ref["Idp"]), "pieces do not hold every particle exactly once"
    worst = float(np.max(np.abs(pos[o] - ref["pos"])))
    print(f"[3] two slab ranks, pieces of PART 1 after 10 steps: max |dpos| vs single GPU = {worst:.3e} m")
    return worst < 1e-6


if __name__ == "__main__":
    ok = True
    with tempfile.TemporaryDirectory() as td:
        for f in (check_mdbc_restart, check_runner_one_gpu, check_pieces_two_ranks):
            try:
                good = f(td)
            except Exception as e:  # noqa: BLE001
                good = False
                print(f"{f.__name__}: {type(e).__name__}: {e}")
            print(f"{f.__name__}: {'ok' if good else 'FAILED'}")
            ok = ok and good
    sys.exit(0 if ok else 1)
