"""Case writer: emits ``<name>.xml`` + ``<name>.bi4`` in the format GenCase produces.

GenCase is closed source and has no Linux binary in the reference tree, so both
the parity oracle (the compiled reference CPU solver) and this repository's
solver consume cases produced here. Only what the solver reads is written
(file:line relative to /root/reference/src/source):

* ``case.execution.constants``  -> JCaseCtes::ReadXmlRun    (JCaseCtes.cpp:208-222)
* ``case.execution.parameters`` -> JSph::LoadConfigParameters (JSph.cpp:663-854)
* ``case.execution.particles``  -> JCaseParts::ReadXml       (JCaseParts.cpp:540-567,79-84,274-297)
* ``case.execution.motion``     -> JMotion::ReadXml           (JMotion.cpp:692-903)
* ``<name>.bi4``                -> JPartsLoad4::LoadParticles (JPartsLoad4.cpp:164-290)

Geometry follows examples/main/01_DamBreak/CaseDambreak_Def.xml:28-38: a tank
(bottom + four walls, open top) on a cubic lattice of spacing dp with a fluid
column against the x=0 wall.  Particle order / ids: fixed, moving, floating,
fluid (JSph.cpp:1030-1036).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import numpy as np

from . import bi4


@dataclass
class DamBreakSpec:
    dp: float = 0.02
    tank: tuple = (1.6, 0.67, 0.4)          # inner size L, W, H (lattice rounded)
    fluid: tuple = (0.4, 0.67, 0.3)         # column size lx, ly, lz from the x=0,y=0,z=0 corner
    layers: int = 1                         # boundary thickness in particles (>=1; mDBC wants >=4)
    normals: bool = False                   # write BoundNor (boundary->interface vectors) for mDBC
    piston: bool = False                    # moving plate (mkbound=1) near x=0 driven by mvrectsinu
    piston_ampl: float = 0.02
    piston_freq: float = 1.0
    motion_xml: str = ""                    # replaces the default mvrectsinu of the piston: raw xml inside <objreal ref="0">
    floating_box: tuple | None = None       # (x0,y0,z0,sx,sy,sz) box of floating particles, relative density
    floating_rel_density: float = 0.5
    gravity: tuple = (0.0, 0.0, -9.81)
    rhop0: float = 1000.0
    gamma: float = 7.0
    coefsound: float = 20.0
    coefh: float = 1.0
    cfl: float = 0.2
    hydrostatic: bool = True                # rhopgradient=2 (water column)
    open_axes: str = ""                     # "y", "xy": no tank walls across these axes (periodic cases); the lattice then
                                            # covers [0, n-1] along them so that the period is n*dp
    fluid_vel: tuple = (0.0, 0.0, 0.0)      # initial velocity of the fluid particles
    two_d: bool = False                     # 2-D case (<data2d value="true">): one lattice plane at y=0, no walls across y,
                                            # h = coefh*sqrt(2)*dp, particle mass = rho0*dp^2 (GenCase's 2-D conventions)
    params: dict = field(default_factory=dict)   # overrides for <parameters>
    special_xml: str = ""                   # raw xml placed inside <special>


DEFAULT_PARAMS = {
    "SavePosDouble": 1, "StepAlgorithm": 2, "VerletSteps": 40, "Kernel": 2, "ViscoTreatment": 1,
    "Visco": 0.1, "ViscoBoundFactor": 1, "Boundary": 1, "SlipMode": 1, "DensityDT": 2,
    "DensityDTvalue": 0.1, "Shifting": 0, "RigidAlgorithm": 1, "FtPause": 0.0, "CoefDtMin": 0.05,
    "DtIni": 0, "DtMin": 0, "DtFixed": 0, "DtAllParticles": 0, "TimeMax": 0.1, "TimeOut": 0.01,
    "RhopOutMin": 700, "RhopOutMax": 1300, "MinFluidStop": 0,
}


def _lattice(i0, i1, j0, j1, k0, k1):
    """Integer lattice points i in [i0,i1], j in [j0,j1], k in [k0,k1], x fastest."""
    if i1 < i0 or j1 < j0 or k1 < k0:
        return np.zeros((0, 3), dtype=np.int64)
    k, j, i = np.meshgrid(np.arange(k0, k1 + 1), np.arange(j0, j1 + 1), np.arange(i0, i1 + 1),
                          indexing="ij")
    return np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1).astype(np.int64)


def build_particles(spec: DamBreakSpec) -> dict:
    """Lattice particles of the dam-break tank; returns arrays in file order."""
    dp = spec.dp
    nx = int(round(spec.tank[0] / dp))
    ny = int(round(spec.tank[1] / dp))
    nz = int(round(spec.tank[2] / dp))
    nl = max(1, int(spec.layers))
    # -- boundary shell: lattice box [-(nl-1) .. n+(nl-1)] minus the interior, open top
    lo = -(nl - 1)
    ox, oy = "x" in spec.open_axes, "y" in spec.open_axes
    if spec.two_d:
        oy, ny = True, 1                    # the single plane j=0
    full = _lattice(0 if ox else lo, nx - 1 if ox else nx + nl - 1, 0 if oy else lo, ny - 1 if oy else ny + nl - 1, lo, nz)
    i, j, k = full[:, 0], full[:, 1], full[:, 2]
    interior = (ox | ((i >= 1) & (i <= nx - 1))) & (oy | ((j >= 1) & (j <= ny - 1))) & (k >= 1)
    fixed = full[~interior]
    # -- fluid column
    fx = min(nx - 1, int(math.floor(spec.fluid[0] / dp + 1e-9)))
    fy = 0 if spec.two_d else min(ny - 1, int(math.floor(spec.fluid[1] / dp + 1e-9)))
    fz = min(nz - 1, int(math.floor(spec.fluid[2] / dp + 1e-9)))
    fluid = _lattice(0 if ox else 1, fx, 0 if oy else 1, fy, 1, fz)
    moving = np.zeros((0, 3), dtype=np.int64)
    floating = np.zeros((0, 3), dtype=np.int64)
    if spec.piston:
        # plate two lattice planes thick at i=2,3 (between wall and fluid); fluid starts at i=5
        moving = _lattice(2, 3, 1, ny - 1, 1, nz)
        fluid = fluid[fluid[:, 0] >= 5]
    if spec.floating_box is not None:
        x0, y0, z0, sx, sy, sz = spec.floating_box
        a = [int(round(v / dp)) for v in (x0, y0, z0)]
        b = [int(round((u + v) / dp)) for u, v in ((x0, sx), (y0, sy), (z0, sz))]
        floating = _lattice(a[0], b[0], a[1], b[1], a[2], b[2])
        # remove fluid overlapping the box (keep one dp gap)
        fi, fj, fk = fluid[:, 0], fluid[:, 1], fluid[:, 2]
        inside = ((fi >= a[0] - 1) & (fi <= b[0] + 1) & (fj >= a[1] - 1) & (fj <= b[1] + 1)
                  & (fk >= a[2] - 1) & (fk <= b[2] + 1))
        fluid = fluid[~inside]
    out = {"nx": nx, "ny": ny, "nz": nz, "nl": nl, "fixed": fixed, "moving": moving,
           "floating": floating, "fluid": fluid, "fz_top": fz}
    return out


def _normals(spec: DamBreakSpec, ijk: np.ndarray, nx: int, ny: int, nz: int) -> np.ndarray:
    """Boundary->interface vectors for the tank shell (mDBC; JSph.cpp:1397-1417).

    The interface lies dp/2 inside the innermost boundary layer: planes x=dp/2,
    x=(nx-1/2)dp, y=..., z=dp/2. The normal of a particle is the vector to the closest
    point of the fluid-side box [dp/2, (n-1/2)dp]^2 x [dp/2, inf).
    """
    dp = spec.dp
    p = ijk.astype(np.float64) * dp
    lo = np.array([0.5 * dp, 0.5 * dp, 0.5 * dp])
    hi = np.array([(nx - 0.5) * dp, (ny - 0.5) * dp, np.inf])
    for k, ax in enumerate("xy"):           # no wall across an open (periodic) axis
        if ax in spec.open_axes:
            lo[k], hi[k] = -np.inf, np.inf
    q = np.minimum(np.maximum(p, lo), hi)
    return (q - p).astype(np.float32)


def constants(spec: DamBreakSpec, parts: dict) -> dict:
    dp = spec.dp
    g = abs(spec.gravity[2])
    hswl = parts["fz_top"] * dp  # fluid top (z of highest fluid particle) measured from z=0... GenCase uses fluid height
    h = spec.coefh * math.sqrt((2.0 if spec.two_d else 3.0) * dp * dp)
    cs0 = spec.coefsound * math.sqrt(g * hswl)
    b = cs0 * cs0 * spec.rhop0 / spec.gamma
    mass = spec.rhop0 * dp ** (2 if spec.two_d else 3)
    return {"dp": dp, "h": h, "b": b, "cs0": cs0, "hswl": hswl, "massfluid": mass, "massbound": mass}


def write_case(dirpath: str, name: str, spec: DamBreakSpec) -> dict:
    """Write ``dirpath/name.xml`` and ``dirpath/name.bi4``; returns a summary dict."""
    os.makedirs(dirpath, exist_ok=True)
    parts = build_particles(spec)
    c = constants(spec, parts)
    dp = spec.dp
    blocks = [parts["fixed"], parts["moving"], parts["floating"], parts["fluid"]]
    nfixed, nmoving, nfloat, nfluid = (int(b.shape[0]) for b in blocks)
    ijk = np.concatenate(blocks, axis=0)
    n = int(ijk.shape[0])
    nbound = nfixed + nmoving + nfloat
    pos = ijk.astype(np.float64) * dp
    vel = np.zeros((n, 3), dtype=np.float32)
    vel[nbound:] = np.asarray(spec.fluid_vel, dtype=np.float32)
    rhop = np.full(n, spec.rhop0, dtype=np.float32)
    if spec.hydrostatic:
        # rhopgradient=2: rho = rho0*(1+rho0*g*(hswl-z)/B)^(1/gamma), fluid only, column top = hswl
        z = pos[nbound:, 2]
        depth = np.maximum(c["hswl"] - z, 0.0)
        rhop[nbound:] = (spec.rhop0 * (1.0 + spec.rhop0 * abs(spec.gravity[2]) * depth / c["b"])
                         ** (1.0 / spec.gamma)).astype(np.float32)
    idp = np.arange(n, dtype=np.uint32)
    boundnor = None
    if spec.normals:
        boundnor = np.zeros((nbound, 3), dtype=np.float32)
        boundnor[:nfixed] = _normals(spec, parts["fixed"], parts["nx"], parts["ny"], parts["nz"])
        if nmoving:
            # piston plate (lattice planes i=2,3, fluid from i=5): interface dp/2 in front of the plane i=3
            boundnor[nfixed:nfixed + nmoving, 0] = (3.5 * dp - pos[nfixed:nfixed + nmoving, 0]).astype(np.float32)

    posmin, posmax = pos.min(axis=0), pos.max(axis=0)
    root = bi4.make_case_item(name, np_total=n, nfixed=nfixed, nmoving=nmoving, nfloat=nfloat,
                              nfluid=nfluid, posmin=posmin, posmax=posmax, dp=dp, h=c["h"], b=c["b"],
                              rhop0=spec.rhop0, gamma=spec.gamma, massbound=c["massbound"],
                              massfluid=c["massfluid"], data2d=bool(spec.two_d), data2dposy=0.0)
    bi4.add_part(root, 0, timestep=0.0, idp=idp, pos=pos, vel=vel, rhop=rhop, posdouble=True,
                 boundnor=boundnor)
    bi4.save(os.path.join(dirpath, name + ".bi4"), root)

    # ------------------------------------------------------------------ XML
    prm = dict(DEFAULT_PARAMS)
    prm.update(spec.params)
    x = []
    x.append('<?xml version="1.0" encoding="UTF-8" ?>')
    x.append('<case app="b200sph-casegen v1" date="00-00-0000 00:00:00">')
    x.append('  <execution>')
    x.append('    <special>')
    if spec.special_xml:
        x.append(spec.special_xml)
    x.append('    </special>')
    x.append('    <parameters>')
    for k, v in prm.items():
        x.append(f'      <parameter key="{k}" value="{v}" />')
    x.append('      <simulationdomain>')
    x.append('        <posmin x="default" y="default" z="default" />')
    x.append('        <posmax x="default" y="default" z="default + 50%" />')
    x.append('      </simulationdomain>')
    x.append('    </parameters>')
    x.append(f'    <particles np="{n}" nb="{nbound}" nbf="{nfixed}" mkboundfirst="11" mkfluidfirst="1">')
    begin = 0
    mkb = 0
    if nfixed:
        x.append(f'      <fixed mkbound="{mkb}" mk="{11 + mkb}" begin="{begin}" count="{nfixed}" />')
        begin += nfixed
        mkb += 1
    if nmoving:
        x.append(f'      <moving mkbound="{mkb}" mk="{11 + mkb}" begin="{begin}" count="{nmoving}" refmotion="0" />')
        begin += nmoving
        mkb += 1
    ft = None
    if nfloat:
        pf = pos[nfixed + nmoving:nbound]
        masspart = c["massbound"] * spec.floating_rel_density
        massbody = masspart * nfloat
        center = np.array([float(v) for v in pf.mean(axis=0)])
        d = pf - center
        center = [float(v) for v in center]
        # inertia tensor of the particle cloud (GenCase convention: sum of point masses)
        ixx = masspart * float(np.sum(d[:, 1] ** 2 + d[:, 2] ** 2))
        iyy = masspart * float(np.sum(d[:, 0] ** 2 + d[:, 2] ** 2))
        izz = masspart * float(np.sum(d[:, 0] ** 2 + d[:, 1] ** 2))
        ixy = -masspart * float(np.sum(d[:, 0] * d[:, 1]))
        ixz = -masspart * float(np.sum(d[:, 0] * d[:, 2]))
        iyz = -masspart * float(np.sum(d[:, 1] * d[:, 2]))
        ft = {"massbody": massbody, "masspart": masspart, "center": list(center),
              "inertia": [[ixx, ixy, ixz], [ixy, iyy, iyz], [ixz, iyz, izz]]}
        x.append(f'      <floating mkbound="{mkb}" mk="{11 + mkb}" begin="{begin}" count="{nfloat}">')
        x.append(f'        <massbody value="{massbody!r}" />')
        x.append(f'        <masspart value="{masspart!r}" />')
        x.append(f'        <center x="{center[0]!r}" y="{center[1]!r}" z="{center[2]!r}" />')
        x.append('        <inertia>')
        for r, row in enumerate(ft["inertia"]):
            x.append(f'          <values v{r + 1}1="{row[0]!r}" v{r + 1}2="{row[1]!r}" v{r + 1}3="{row[2]!r}" />')
        x.append('        </inertia>')
        x.append('      </floating>')
        begin += nfloat
        mkb += 1
    x.append(f'      <fluid mkfluid="0" mk="1" begin="{begin}" count="{nfluid}" />')
    x.append('    </particles>')
    x.append('    <constants>')
    x.append('      <data2d value="true" />' if spec.two_d else '      <data2d value="false" />')
    if spec.two_d:
        x.append('      <data2dposy value="0" />')
    x.append(f'      <gravity x="{spec.gravity[0]!r}" y="{spec.gravity[1]!r}" z="{spec.gravity[2]!r}" />')
    x.append(f'      <cflnumber value="{spec.cfl!r}" />')
    x.append(f'      <gamma value="{spec.gamma!r}" />')
    x.append(f'      <rhop0 value="{spec.rhop0!r}" />')
    x.append(f'      <dp value="{dp!r}" />')
    x.append(f'      <h value="{c["h"]:.10E}" />')
    x.append(f'      <b value="{c["b"]:.10E}" />')
    x.append(f'      <massbound value="{c["massbound"]:.10E}" />')
    x.append(f'      <massfluid value="{c["massfluid"]:.10E}" />')
    x.append('    </constants>')
    if nmoving:
        x.append('    <motion>')
        x.append('      <objreal ref="0">')
        if spec.motion_xml:
            x.append(spec.motion_xml)
        else:
            x.append('        <begin mov="1" start="0" />')
            x.append('        <mvrectsinu id="1" duration="1000" anglesunits="radians">')
            x.append(f'          <freq x="{spec.piston_freq!r}" y="0" z="0" />')
            x.append(f'          <ampl x="{spec.piston_ampl!r}" y="0" z="0" />')
            x.append('          <phase x="0" y="0" z="0" />')
            x.append('        </mvrectsinu>')
        x.append('      </objreal>')
        x.append('    </motion>')
    x.append('  </execution>')
    x.append('</case>')
    with open(os.path.join(dirpath, name + ".xml"), "w") as f:
        f.write("\n".join(x) + "\n")
    return {"np": n, "nfixed": nfixed, "nmoving": nmoving, "nfloat": nfloat, "nfluid": nfluid,
            "constants": c, "floating": ft, "xml": os.path.join(dirpath, name + ".xml"),
            "bi4": os.path.join(dirpath, name + ".bi4")}


def dp_for_target(np_target: int, tank=(1.6, 0.67, 0.4), fluid=(0.4, 0.67, 0.3), layers: int = 1) -> float:
    """dp such that fluid+boundary particle count is close to ``np_target`` (bisection on dp)."""
    def count(dp):
        nx, ny, nz = (int(round(t / dp)) for t in tank)
        nl = layers
        tot = (nx + 2 * nl - 1) * (ny + 2 * nl - 1) * (nz + nl)
        inner = max(nx - 1, 0) * max(ny - 1, 0) * nz
        fx = min(nx - 1, int(math.floor(fluid[0] / dp + 1e-9)))
        fy = min(ny - 1, int(math.floor(fluid[1] / dp + 1e-9)))
        fz = min(nz - 1, int(math.floor(fluid[2] / dp + 1e-9)))
        return (tot - inner) + fx * fy * fz
    lo, hi = 1e-4, 0.2
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if count(mid) > np_target:
            lo = mid
        else:
            hi = mid
    return hi
