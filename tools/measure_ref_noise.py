#!/usr/bin/env python
"""Reference-vs-reference noise: how far two builds of the UNMODIFIED reference CPU solver are apart on the parity fixtures.

    make -C oracle -j6 ref refnofm            (both probes: the shipped flags, and the same without -ffast-math)
    python tools/measure_ref_noise.py          -> profiles/ref_noise_r02.json

For every fixture of tools/make_golden.py (and the 170k-particle bench configuration of tools/make_golden_large.py) the case is
regenerated, stepped by oracle/_ref/dsph_ref_probe_nofm (g++ -O3 without -ffast-math), and its dumps are compared with the
committed golden dumps of the shipped -ffast-math build -- same metric as tests/test_gpu_parity.py (max |a-b| / max|b|; ar with
the floor rho0*cs0/h*1e-4; positions absolute in m).  The two builds run the same source on the same inputs, so this is the floor
below which "agrees with the reference" has no meaning: the per-field tolerances of the GPU parity tests are set to
max(3 x measured GPU error, this noise) (tests/tolerances.py)."""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from dualsphysics_b200 import b2sd, casegen  # noqa: E402
from oracle import oracle as orc  # noqa: E402
import make_golden  # noqa: E402
import make_golden_large  # noqa: E402

NOFM = os.path.join(orc.REF_DIR, "dsph_ref_probe_nofm")


def rel(a, b, floor=None):
    a, b = np.asarray(a, np.float64).reshape(-1), np.asarray(b, np.float64).reshape(-1)
    sc = np.max(np.abs(b)) if floor is None else max(np.max(np.abs(b)), floor)
    return float(np.max(np.abs(a - b)) / sc) if sc else float(np.max(np.abs(a - b)))


def compare(gn, gf, floor):
    """gn: dump of the no-fast-math build, gf: dump of the shipped build (same particle order is checked first)."""
    e = {}
    if "idp" in gn and "idp" in gf and not np.array_equal(gn["idp"], gf["idp"]):
        return {"order_differs": True}
    n = int(gf["np"])
    own = (np.asarray(gf["code"]) & 0xe000) == 0 if "code" in gf else np.ones(n, bool)
    if "ace" in gf:
        e["press"] = rel(gn["press"], gf["press"])
        e["ace"] = rel(np.asarray(gn["ace"]).reshape(-1, 3)[own], np.asarray(gf["ace"]).reshape(-1, 3)[own])
        e["ar"] = rel(np.asarray(gn["ar"])[own], np.asarray(gf["ar"])[own], floor)
        for k in ("acemax", "viscdtmax", "velmax"):
            e[k] = abs(float(gn[k]) - float(gf[k])) / max(abs(float(gf[k])), 1e-6 if k != "acemax" else 1e-30)
        key = "dt_p" if "dt_p" in gf else "dt"
        if key in gf:
            e["dt"] = abs(float(gn[key]) - float(gf[key])) / float(gf[key])
    if "stepdt" in gf:
        e["pos_abs_m"] = float(np.max(np.abs(np.asarray(gn["pos"], np.float64) - np.asarray(gf["pos"], np.float64))))
        e["velrho"] = rel(gn["velrho"], gf["velrho"])
        e["stepdt"] = abs(float(gn["stepdt"]) - float(gf["stepdt"])) / float(gf["stepdt"])
    return e


def main(names=None):
    """names: only these fixtures; their records replace the old ones in profiles/ref_noise_r02.json, the others are kept."""
    if not os.path.exists(NOFM):
        raise SystemExit("build it first: make -C oracle -j6 refnofm")
    orc_probe = orc.REF_PROBE
    records = []
    jobs = [(name, kw, opts, nsteps, keep, None) for name, (kw, opts, nsteps, keep) in make_golden.specs().items()]
    jobs.append(("c1_170k_bench_half", dict(dp=0.0085), make_golden_large.OPTS, 10, make_golden_large.KEEP, "fresh"))
    for name, kw, opts, nsteps, keep, mode in jobs:
        if names and name not in names:
            continue
        with tempfile.TemporaryDirectory() as td:
            casegen.write_case(td, "Case", casegen.DamBreakSpec(**kw))
            xml = open(os.path.join(td, "Case.xml")).read()
            import re
            h = float(re.search(r'<h value="([^"]+)"', xml).group(1))
            rho0 = float(re.search(r'<rhop0 value="([^"]+)"', xml).group(1))
            cteb = float(re.search(r'<b value="([^"]+)"', xml).group(1))
            gamma = float(re.search(r'<gamma value="([^"]+)"', xml).group(1))
            cs0 = (cteb * gamma / rho0) ** 0.5          # B = cs0^2*rho0/gamma (JSph.cpp:1507)
            threads = 1 if mode is None else (os.cpu_count() or 1)
            orc.REF_PROBE = NOFM
            pn = orc.run_ref_probe(os.path.join(td, "Case"), os.path.join(td, "outn"), nsteps, options=opts, threads=threads)
            orc.REF_PROBE = orc_probe
            if mode == "fresh":   # the committed large fixture is subsampled: run the shipped build again for the full fields
                pf = orc.run_ref_probe(os.path.join(td, "Case"), os.path.join(td, "outf"), nsteps, options=opts, threads=threads)
            else:
                pf = os.path.join(ROOT, "tests", "golden", name)
            for tag in keep:
                fn, ff = os.path.join(pn, tag + ".b2sd"), os.path.join(pf, tag + ".b2sd")
                if not (os.path.exists(fn) and os.path.exists(ff)):
                    continue
                gn, gf = b2sd.load(fn), b2sd.load(ff)
                if "ace" not in gf and "stepdt" not in gf:
                    continue
                e = compare(gn, gf, rho0 * cs0 / h * 1e-4)
                records.append({"key": name, "dump": tag, **e})
                print(name, tag, json.dumps(e), flush=True)
    outpath = os.path.join(ROOT, "profiles", "ref_noise_r02.json")
    if names and os.path.exists(outpath):
        records = [r for r in json.load(open(outpath))["records"] if r["key"] not in names] + records
    agg = {}
    for r in records:
        for k, v in r.items():
            if isinstance(v, float):
                agg[k] = max(agg.get(k, 0.0), v)
    out = {"what": "max difference per field between two builds of the unmodified reference CPU solver (g++ -O3 -fopenmp with and "
                   "without -ffast-math) on the parity fixtures; relative to max|field| (ar: floor rho0*cs0/h*1e-4; pos absolute in m); "
                   "tools/measure_ref_noise.py", "max_over_fixtures": agg, "records": records}
    with open(outpath, "w") as f:
        json.dump(out, f, indent=1)
    print("max over fixtures:", json.dumps(agg))


if __name__ == "__main__":
    main(sys.argv[1:])
