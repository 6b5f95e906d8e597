"""Generates tests/golden/ref_v1.json from oracle/_ref/drrt_ref -- the
REFERENCE'S OWN sources (kdtree.cpp, ghostpoint.cpp, jlist.cpp, dubinsedge.cpp,
obstacle.cpp, distancefunctions.cpp) compiled in this container by
`make -C oracle ref`.  /root/reference does not exist on the GPU box, so the
reference's outputs on these seeded inputs are committed as a fixture; the
oracle (and through it the CUDA path) is checked against them everywhere.

    make -C oracle ref && python tests/golden/make_ref_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from drrt_b200 import synth  # noqa: E402
from oracle import pyref  # noqa: E402

W = synth.TWO_PI


def inputs():
    pos = synth.box_points(6000, 0xB1)
    pos[0] = [0.0, 0.0, -3 * synth.PI / 4]       # smalltest.cpp start pose: negative heading
    q = synth.box_points(60, 0xB2)
    q[:10, 2] = np.linspace(0.0, 0.25, 10)
    q[10:20, 2] = W - np.linspace(0.0, 0.25, 10)
    q[20] = pos[0]
    rad = np.linspace(0.4, 3.0, 60)
    o = synth.obstacles(32, 0xB3)
    o["kind"][:4] = [1, 2, 5, 4]
    o["used"][4] = 0
    o["life_span"][5] = 0.0
    s, e = synth.random_edges(500, 0xB4, max_len=3.0)
    s[:40], e[:40] = synth.random_edges(40, 0xB5, max_len=0.5)
    e[:10, 2] = s[:10, 2]                          # "0s0" candidates (dist <= 2, equal headings)
    pts = synth.box_points(400, 0xB6)
    grid = np.stack([a.ravel() for a in np.meshgrid(np.arange(12.0), np.arange(12.0), indexing="ij")]
                    + [np.zeros(144)], 1)
    return dict(pos=pos, q=q, rad=rad, o=o, s=s, e=e, pts=pts, grid=grid)


def hexes(a):
    return [float(x).hex() for x in np.asarray(a, dtype=np.float64).ravel()]


def main():
    assert pyref.available(), "build oracle/_ref first: make -C oracle ref"
    I = inputs()
    out = {}
    st = pyref.structure(I["pos"])
    out["structure"] = {k: v.tolist() for k, v in st.items()}
    for name, wrap, metric in (("wrap_r2s1", True, 0), ("nowrap_r2s1", False, 0), ("wrap_euclid", True, 1)):
        off, ids, dist = pyref.range_batch(I["pos"], I["q"], I["rad"], metric=metric, wrap=wrap)
        nid, nd = pyref.nearest_batch(I["pos"], I["q"], metric=metric, wrap=wrap)
        out["range_" + name] = dict(off=off.tolist(), ids=ids.tolist(), dist=hexes(dist))
        out["nearest_" + name] = dict(ids=nid.tolist(), dist=hexes(nd))
    # Theta*'s lattice: distances land exactly on the radius (thetastar.cpp:80-110, 152)
    off, ids, dist = pyref.range_batch(I["grid"], I["grid"][::5], 2.0, metric=1, wrap=True)
    out["range_lattice"] = dict(off=off.tolist(), ids=ids.tolist(), dist=hexes(dist))
    cnt, g = pyref.ghosts(I["q"], 0.6)
    out["ghosts"] = dict(counts=cnt.tolist(), points=hexes(g))
    r = pyref.edges(I["s"], I["e"], I["o"], want_traj=True)
    out["edges"] = dict(type=r["type"], dist=hexes(r["dist"]), npts=r["npts"].tolist(),
                        verdict=r["verdict"].tolist(),
                        xy_first=[hexes(x[:3]) for x in r["xy"]], xy_last=[hexes(x[-2:]) for x in r["xy"]])
    out["points"] = pyref.points(I["pts"], I["o"]).tolist()
    rng = np.random.default_rng(0xB7)
    seg = rng.uniform(0, 4, (300, 8))
    seg[:20, 2] = seg[:20, 0]                     # vertical P (|dx| < 1e-6 branch)
    seg[20:40, 6] = seg[20:40, 4]                 # vertical Q
    sd, pd = pyref.geom(seg[:, 0:2], seg[:, 2:4], seg[:, 4:6], seg[:, 6:8])
    out["geom"] = dict(segs=hexes(seg), seg_dist=hexes(sd), pt_dist=hexes(pd))
    path = os.path.join(ROOT, "tests", "golden", "ref_v1.json")
    with open(path, "w") as fh:
        json.dump(out, fh)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
