"""PartMotionRef.ibi4 (the file the post-processing tools rebuild body motion from): our writer, fed with the boundary positions of
the reference's own PART files, against the file the reference CPU build wrote for the same run (tests/golden/ref_part_motion/,
tools/make_ref_part_motion.py).  Host code only: same head (names, types), the same three reference particles per body -- which
pins the corner search, its tie-breaking and the reference's argument order in the line-distance step -- and the same rows."""
import json
import os

import numpy as np

import helpers as H
from dualsphysics_b200 import bi4, lib as L

REF = os.path.join(H.GOLDEN, "ref_part_motion")


def test_motionref_file_matches_reference(tmp_path, built_lib):
    ref = bi4.load_list(os.path.join(REF, "PartMotionRef.ibi4"))
    z = np.load(os.path.join(REF, "bound.npz"))
    nparts = int(z["nparts"])
    hd = ref[0]
    nmk, nmv = hd.get("MkCount"), hd.get("MkMovingCount")
    # blocks of the case (fixed | piston | floating box): begin/count from the ids of the reference particles is not enough, so
    # regenerate the case the fixture was made from and read its MK blocks
    H.write_case(tmp_path, **json.load(open(os.path.join(REF, "fixture_motionref.json")))["spec"])
    _, aux = H.params_from_case(str(tmp_path / "Case"))
    bodies = [(tag == "floating", int(el.get("mkbound")), begin, count) for tag, begin, count, el in aux["blocks"] if tag in ("moving", "floating")]
    assert len(bodies) == nmk and sum(1 for b in bodies if not b[0]) == nmv
    idp = z["idp0"]
    pos = np.stack([z["pos%d" % k] for k in range(nparts)])
    for k in range(nparts):
        assert np.array_equal(z["idp%d" % k], idp)
    out = tmp_path / "data"
    out.mkdir()
    L.motionref_write(str(out), bodies, hd.get("TimeOut"), hd.get("MkBoundFirst"), idp, pos,
                      [float(z["time%d" % k]) for k in range(nparts)], [int(z["step%d" % k]) for k in range(nparts)])
    ours = bi4.load_list(str(out / "PartMotionRef.ibi4"))
    assert len(ours) == len(ref) == nparts + 1
    for a, b in zip(ours, ref):
        assert a.name == b.name and a.hide == b.hide
        assert [(n, t) for n, t, _ in a.values] == [(n, t) for n, t, _ in b.values], a.name
        assert [(n, t, len(v)) for n, t, v in a.arrays] == [(n, t, len(v)) for n, t, v in b.arrays], a.name
        for (n, t, va), (_, _, vb) in zip(a.values, b.values):
            if n != "AppName":
                assert va == vb, (a.name, n, va, vb)
        for (n, _, va), (_, _, vb) in zip(a.arrays, b.arrays):
            va, vb = np.asarray(va), np.asarray(vb)
            if n == "Dis":      # sqrt(...)*1.1: the reference build is -ffast-math, one ulp apart
                assert np.allclose(va, vb, rtol=1e-14, atol=0), (a.name, n)
            else:               # ids, selected particles, their positions at every PART: identical
                assert np.array_equal(va, vb), (a.name, n)
