#!/usr/bin/env python
"""Generates tests/golden_large/* from the COMPILED REFERENCE (oracle/_ref/dsph_ref_probe) at BASELINE sizes.

    python tools/make_golden_large.py            (build container; needs `make -C oracle ref`)

The fixtures of tests/golden are ~1.8k particles.  These are the bench configuration (Symplectic, Wendland, artificial viscosity,
DDT2, DBC, -cellmode:half) at BASELINE configs[0] size (dp 0.0085, ~170k particles) and at ~1M particles, 10 steps of the unmodified
DualSPHysics v5.4 CPU solver.  A full dump is ~100 MB, so a fixture keeps
  * the scalars (counts, cell grid, maxima, dt),
  * a 64-bit position-weighted checksum of every integer table (cellpart, sortpart, begincell, idp, dcell, code): the test
    recomputes it over the whole array, so the tables are still compared bit for bit,
  * the float fields (ace, ar, press, pos, velrho) of a fixed subsample: every particle whose id is a multiple of `stride`.
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dualsphysics_b200 import b2sd, casegen  # noqa: E402
from oracle import oracle as orc  # noqa: E402

OPTS = ["-symplectic", "-wendland", "-viscoart:0.1", "-ddt:2", "-ddtvalue:0.1", "-dbc", "-cellmode:half"]
FIXTURES = {   # name -> (np target or None for the shipped dp, dp, id stride)
    "c1_170k_bench_half": (None, 0.0085, 29),
    "c2_1m_bench_half": (1000000, None, 173),
}
KEEP = ["s0000_pred", "s0009_pred", "s0009_end"]
INT_TABLES = ["cellpart", "sortpart", "begincell", "idp", "dcell", "code"]


def table_hash(a) -> int:
    """sum_i a[i]*(i*0x9E3779B1+1) mod 2^64: order-sensitive, computable with numpy on any box."""
    a = np.ascontiguousarray(a).reshape(-1).astype(np.uint64)
    w = np.arange(a.size, dtype=np.uint64) * np.uint64(0x9E3779B1) + np.uint64(1)
    with np.errstate(over="ignore"):
        return int(np.sum(a * w, dtype=np.uint64))


def reduce_dump(g: dict, stride: int) -> dict:
    out = {}
    n = int(g["np"])
    idp = np.asarray(g["idp"]) if "idp" in g else None
    sel = np.nonzero(idp % stride == 0)[0] if idp is not None else None
    for k, v in g.items():
        a = np.asarray(v)
        if a.ndim == 0 or a.size == 1:
            out[k] = v
        elif k in INT_TABLES:
            out["hash_" + k] = np.array([table_hash(a) & 0xffffffff, table_hash(a) >> 32], dtype=np.uint32)
            out["len_" + k] = np.uint32(a.size)
            if k == "idp":
                out["sub_idp"] = a[sel]
        elif sel is not None and a.size % n == 0 and a.size >= n:
            out["sub_" + k] = a.reshape(n, -1)[sel].reshape(-1)
    return out


def main(names=None):
    base = os.path.join(ROOT, "tests", "golden_large")
    os.makedirs(base, exist_ok=True)
    for name, (npt, dp, stride) in FIXTURES.items():
        if names and name not in names:
            continue
        dpv = dp if dp else casegen.dp_for_target(npt)
        with tempfile.TemporaryDirectory() as td:
            casegen.write_case(td, "Case", casegen.DamBreakSpec(dp=dpv))
            probe = orc.run_ref_probe(os.path.join(td, "Case"), os.path.join(td, "out"), 10, options=OPTS, threads=os.cpu_count())
            dst = os.path.join(base, name)
            os.makedirs(dst, exist_ok=True)
            for k in KEEP:
                g = b2sd.load(os.path.join(probe, k + ".b2sd"))
                b2sd.save(os.path.join(dst, k + ".b2sd"), reduce_dump(g, stride))
            with open(os.path.join(dst, "fixture.json"), "w") as f:
                json.dump({"spec": {"dp": dpv}, "ref_options": OPTS, "nsteps": 10, "stride": stride,
                           "reference": "DualSPHysics v5.4.355 CPU, gcc -O3 -ffast-math -fopenmp"}, f, indent=1)
            print(name, "->", dst, sum(os.path.getsize(os.path.join(dst, f)) for f in os.listdir(dst)) // 1024, "KiB")


if __name__ == "__main__":
    main(sys.argv[1:])
