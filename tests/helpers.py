"""Shared helpers for the test-suite: case generation, host-side parameter loading, comparisons."""
from __future__ import annotations

import ctypes as C
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import dualsphysics_b200 as dsb  # noqa: E402
from dualsphysics_b200 import bi4, b2sd, casegen  # noqa: E402
from oracle import oracle as orc  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def have_ref() -> bool:
    return os.path.exists(orc.REF_PROBE)


def have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def write_case(dirpath, name="Case", **kw):
    spec = casegen.DamBreakSpec(**kw)
    return casegen.write_case(str(dirpath), name, spec), spec


def params_from_case(casepath: str, options: dict | None = None):
    """Python twin of the C++ host loader (csrc/host/sim.cpp) for CPU-only tests: Params + particle arrays."""
    root = ET.parse(casepath + ".xml").getroot()
    exe = root.find("execution")
    ct = exe.find("constants")
    prm = {p.get("key"): p.get("value") for p in exe.find("parameters").findall("parameter")}
    if options:
        prm.update({k: str(v) for k, v in options.items()})
    p = dsb.Params()
    g = ct.find("gravity")
    p.gravity[0], p.gravity[1], p.gravity[2] = float(g.get("x")), float(g.get("y")), float(g.get("z"))
    val = lambda n: float(ct.find(n).get("value"))  # noqa: E731
    p.cflnumber, p.gamma, p.rhopzero, p.dp = val("cflnumber"), val("gamma"), val("rhop0"), val("dp")
    p.kernelh, p.cteb, p.massbound, p.massfluid = val("h"), val("b"), val("massbound"), val("massfluid")
    d2 = ct.find("data2d")
    p.simulate2d = 1 if (d2 is not None and d2.get("value", "false").lower() in ("true", "1")) else 0
    p.tkernel = int(prm.get("Kernel", 2))
    p.tstep = int(prm.get("StepAlgorithm", 1))
    p.verletsteps = int(prm.get("VerletSteps", 40))
    p.tvisco = int(prm.get("ViscoTreatment", 1))
    p.visco = float(prm["Visco"])
    p.viscoboundfactor = float(prm.get("ViscoBoundFactor", 1))
    p.tboundary = int(prm.get("Boundary", 1))
    p.slipmode = int(prm.get("SlipMode", 1)) if p.tboundary == 2 else 0
    p.tdensity = int(prm.get("DensityDT", 0))
    p.ddtvalue = float(prm.get("DensityDTvalue", 0.1)) if p.tdensity else 0.0
    p.coefdtmin = float(np.float32(prm.get("CoefDtMin", 0.05)))
    p.dtini = max(0.0, float(prm.get("DtIni", 0)))
    p.dtmin = max(0.0, float(prm.get("DtMin", 0)))
    p.dt_allparticles = 1 if int(prm.get("DtAllParticles", 0)) == 1 else 0
    p.rhopoutmin, p.rhopoutmax = float(prm.get("RhopOutMin", 700)), float(prm.get("RhopOutMax", 1300))
    p.cellmode_half = int(prm.get("_cellmode_half", 0))
    # periodic boundaries (JSph.cpp:815-827)
    px = py = pz = False
    if "XPeriodicIncY" in prm: p.perixinc[1] = float(prm["XPeriodicIncY"]); px = True    # noqa: E701
    if "XPeriodicIncZ" in prm: p.perixinc[2] = float(prm["XPeriodicIncZ"]); px = True    # noqa: E701
    if "YPeriodicIncX" in prm: p.periyinc[0] = float(prm["YPeriodicIncX"]); py = True    # noqa: E701
    if "YPeriodicIncZ" in prm: p.periyinc[2] = float(prm["YPeriodicIncZ"]); py = True    # noqa: E701
    if "ZPeriodicIncX" in prm: p.perizinc[0] = float(prm["ZPeriodicIncX"]); pz = True    # noqa: E701
    if "ZPeriodicIncY" in prm: p.perizinc[1] = float(prm["ZPeriodicIncY"]); pz = True    # noqa: E701
    for key, flags in (("XYPeriodic", (True, True, False)), ("XZPeriodic", (True, False, True)), ("YZPeriodic", (False, True, True))):
        if key in prm:
            px, py, pz = flags
            for k in range(3):
                p.perixinc[k] = p.periyinc[k] = p.perizinc[k] = 0.0
    p.periactive = (1 if px else 0) | (2 if py else 0) | (4 if pz else 0)
    p.shiftmode = int(prm.get("Shifting", 0))
    p.shiftcoef, p.shifttfs = float(prm.get("ShiftCoef", -2)), float(prm.get("ShiftTFS", 0))
    pa = exe.find("particles")
    cnt = {"fixed": 0, "moving": 0, "floating": 0, "fluid": 0}
    blocks = []
    for b in pa:
        if b.tag in cnt:
            cnt[b.tag] += int(b.get("count"))
            blocks.append((b.tag, int(b.get("begin")), int(b.get("count")), b))
    p.casenfixed, p.casenmoving, p.casenfloat, p.casenfluid = cnt["fixed"], cnt["moving"], cnt["floating"], cnt["fluid"]
    p.casenp = sum(cnt.values())
    data = bi4.load_part(casepath + ".bi4")
    pos = data["pos"]
    idp = data["Idp"]
    cmin, cmax = np.array(data["head"]["CasePosMin"]), np.array(data["head"]["CasePosMax"])
    border = np.full(3, float(np.float32(p.kernelh)) * 0.05)
    for k in range(3):
        if (p.periactive >> k) & 1:
            border[k] = p.dp / 2.0          # JPartsLoad4::CalculeLimitsPos (JPartsLoad4.cpp:389-399)
    rmin, rmax = cmin - border, cmax + border
    sd = exe.find("parameters").find("simulationdomain")
    lo, hi = rmin.copy(), rmax.copy()
    for k, ax in enumerate("xyz"):
        for node, ismin in ((sd.find("posmin") if sd is not None else None, True),
                            (sd.find("posmax") if sd is not None else None, False)):
            v = (node.get(ax, "default") if node is not None else "default").replace(" ", "").lower()
            if v == "default":
                continue
            dif = rmax[k] - rmin[k]
            if v.startswith("default"):
                prc = v.endswith("%")
                num = float(v[8:-1] if prc else v[8:])
                inc = dif * num / 100 if prc else num
                if ismin:
                    lo[k] = rmin[k] - inc
                else:
                    hi[k] = rmax[k] + inc
            else:
                if ismin:
                    lo[k] = float(v)
                else:
                    hi[k] = float(v)
    for k in range(3):
        if (p.periactive >> k) & 1:
            lo[k], hi[k] = rmin[k], rmax[k]     # periodic axes keep the particle limits (JSph.cpp:1472-1475)
        p.maprealposmin[k], p.maprealposmax[k] = lo[k], hi[k]
    lib = dsb.load_library()
    r = lib.b2s_config_finalize(C.byref(p))
    assert r == 0, r
    # particle codes (JSphMk): type | index within type
    tcode = {"fixed": 0x0000, "moving": 0x0800, "floating": 0x1000, "fluid": 0x1800}
    idx = {k: 0 for k in tcode}
    codeofid = np.zeros(p.casenp, np.uint16)
    ftmassp = []
    for tag, beg, n, node in blocks:
        codeofid[beg:beg + n] = tcode[tag] | idx[tag]
        idx[tag] += 1
        if tag == "floating":
            ftmassp.append(float(node.find("masspart").get("value")))
    code = codeofid[idp]
    velrho = np.concatenate([data["Vel"], data["Rhop"][:, None]], axis=1).astype(np.float32)
    floatings = []
    for tag, beg, n, node in blocks:
        if tag == "floating":
            cen = np.array([float(node.find("center").get(a)) for a in "xyz"])
            inert = np.array([[float(v.get(f"v{r + 1}{c + 1}")) for c in range(3)] for r, v in enumerate(node.find("inertia").findall("values"))], np.float32)
            floatings.append({"begin": beg, "count": n, "mass": float(node.find("massbody").get("value")), "massp": float(node.find("masspart").get("value")),
                              "center": cen, "angles": np.zeros(3, np.float32), "fvel": np.zeros(3, np.float32), "fomega": np.zeros(3, np.float32),
                              "inertiaini": inert})
    sp = exe.find("special")
    motion = None
    mo = exe.find("motion")
    if mo is not None and mo.find("objreal") is not None:
        obj = mo.find("objreal")
        for tag, beg, n, node in blocks:
            if tag == "moving" and int(node.get("refmotion", -1)) == int(obj.get("ref", 0)):
                motion = (orc.MotionOracle(obj), beg, n)
    arrays = {"idp": idp, "code": code, "pos": pos, "velrho": velrho, "boundnor": data.get("BoundNor"),
              "ftmassp": np.array(ftmassp, np.float32),
              "damping": orc.damping_zones(sp.find("damping") if sp is not None else None), "motion": motion,
              "floatings": floatings, "ftpause": float(prm.get("FtPause", 0)),
              "gauges_xml": sp.find("gauges") if sp is not None else None, "blocks": blocks,
              "timemax": float(prm.get("TimeMax", 0)), "timeout": float(prm.get("TimeOut", 0))}
    return p, arrays


def oracle_sim_from_case(casepath: str, options: dict | None = None) -> "orc.OracleSim":
    p, a = params_from_case(casepath, options)
    cte = orc.cte_from_params(p)
    sim = _oracle_sim(p, a, cte)
    if a["gauges_xml"] is not None:
        sim.gauges = orc.GaugeOracle(a["gauges_xml"], p, a["blocks"], a["timemax"], a["timeout"])
        sim.timemax = a["timemax"]
    return sim


def _oracle_sim(p, a, cte):
    return orc.OracleSim(cte, a["idp"], a["code"], a["pos"], a["velrho"], npb=p.casenpb, dtini=p.dtini,
                         verletsteps=p.verletsteps, ftomassp=a["ftmassp"] if len(a["ftmassp"]) else None,
                         boundnor=None if a["boundnor"] is None else 2.0 * np.asarray(a["boundnor"], np.float32),
                         damping=a["damping"], motion=a["motion"], floatings=a["floatings"], ftpause=a["ftpause"],
                         slipmode=p.slipmode if p.tboundary == 2 else 1)


def rel_err(a, b, floor=None):
    """max |a-b| / max(|b|) -- error relative to the field's magnitude (the tolerance form SURVEY 7(2) asks for)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    scale = np.max(np.abs(b)) if floor is None else max(np.max(np.abs(b)), floor)
    if scale == 0:
        return float(np.max(np.abs(a - b)))
    return float(np.max(np.abs(a - b)) / scale)


def part_observables(datadir, dp):
    """Per PART file: time, dam front (max x of fluid below z=3dp... any fluid), column height at the left wall, kinetic
    energy per unit mass, max |v|, floating-body centre (mean position of its particles)."""
    head = bi4.load(os.path.join(datadir, "Part_Head.ibi4"))
    nfixed, nmoving, nfloat = head.get("CaseNfixed"), head.get("CaseNmoving"), head.get("CaseNfloat")
    nb = nfixed + nmoving + nfloat
    out = []
    k = 0
    while os.path.exists(os.path.join(datadir, f"Part_{k:04d}.bi4")):
        d = bi4.load_part(os.path.join(datadir, f"Part_{k:04d}.bi4"))
        idp, pos, vel = d["Idp"], d["pos"], d["Vel"]
        fl = idp >= nb
        ft = (idp >= nfixed + nmoving) & (idp < nb)
        v2 = np.sum(vel[fl].astype(np.float64) ** 2, axis=1)
        near_wall = fl & (pos[:, 0] < 4 * dp)
        # pressure "gauge": mean Tait pressure of the fluid in the bottom corner of the end wall (x > L-6dp, z < 6dp)
        rho = d["Rhop"].astype(np.float64)
        rho0, gamma, cteb = float(d["head"].get("Rhop0")), float(d["head"].get("Gamma")), float(d["head"].get("B"))
        xwall = float(pos[~fl & (idp < nfixed), 0].max())
        corner = fl & (pos[:, 0] > xwall - 6 * dp) & (pos[:, 2] < 6 * dp)
        press_corner = float(np.mean(cteb * ((rho[corner] / rho0) ** gamma - 1.0))) if corner.any() else 0.0
        rec = {"t": float(d["part"].get("TimeStep")), "front_x": float(pos[fl, 0].max()), "press_corner": press_corner,
               "ncorner": int(corner.sum()),
               "height_wall": float(pos[near_wall, 2].max()) if near_wall.any() else 0.0,
               "ke": float(0.5 * v2.sum()), "vmax": float(np.sqrt(v2.max())), "nfluid": int(fl.sum())}
        if ft.any():
            rec["ft_center"] = [float(x) for x in pos[ft].mean(axis=0)]
        out.append(rec)
        k += 1
    return out


def parity_report(key: str, values: dict) -> None:
    """Appends one measured-error record to the JSON-lines file named by B2S_PARITY_REPORT (the GPU job collects it into
    profiles/parity_r02.json); no-op when the variable is not set."""
    path = os.environ.get("B2S_PARITY_REPORT")
    if not path:
        return
    import json
    with open(path, "a") as f:
        f.write(json.dumps({"key": key, **values}) + "\n")


def table_hash(a) -> int:
    """Checksum of tools/make_golden_large.py: sum_i a[i]*(i*0x9E3779B1+1) mod 2^64."""
    a = np.ascontiguousarray(a).reshape(-1).astype(np.uint64)
    w = np.arange(a.size, dtype=np.uint64) * np.uint64(0x9E3779B1) + np.uint64(1)
    with np.errstate(over="ignore"):
        return int(np.sum(a * w, dtype=np.uint64))
