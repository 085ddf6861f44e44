import json, os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
import dualsphysics_b200 as dsb
from dualsphysics_b200 import b2sd
fx = sys.argv[1]
meta = json.load(open(os.path.join(H.GOLDEN, fx, "fixture.json")))
td = tempfile.mkdtemp()
H.write_case(td, **meta["spec"])
sim = dsb.Simulation(os.path.join(td, "Case"), options=" ".join(meta["ref_options"]))
p = sim.params
g = b2sd.load(os.path.join(H.GOLDEN, fx, "s0000_pred.b2sd"))
inter = 2 if p.tstep == 2 else 1
sim.ctx.interaction(inter)
ar, ace = sim.ctx.array("ar"), sim.ctx.array("ace")
gar, gace = g["ar"], g["ace"].reshape(-1, 3)
npb = int(g["npb"])
d = np.abs(ar - gar)
i = int(np.argmax(d))
print(fx, "np", len(ar), "npb", npb, "max|ar| gpu", np.abs(ar).max(), "ref", np.abs(gar).max(), "maxdiff", d.max(), "at", i, "gpu", ar[i], "ref", gar[i])
print(" fluid diff max", d[npb:].max(), " bound diff max", d[:npb].max())
big = np.nonzero(d > 1e-3 * max(np.abs(gar).max(), 1))[0]
print(" n big diffs", len(big), big[:20], ar[big[:10]], gar[big[:10]])
print(" tdensity", p.tdensity, "ddtkh", p.ddtkh, "tboundary", p.tboundary, "tkernel", p.tkernel)
pos = sim.ctx.array("pos")
print(" pos of worst", pos[i], "code", sim.ctx.array("code")[i])
