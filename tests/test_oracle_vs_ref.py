"""Pins the oracle to the REFERENCE ITSELF.

tests/golden/ref_v1.json holds outputs of the reference's own sources
(kdtree.cpp, ghostpoint.cpp, jlist.cpp, dubinsedge.cpp, obstacle.cpp,
distancefunctions.cpp, compiled unmodified by `make -C oracle ref`; generator:
tests/golden/make_ref_golden.py).  The oracle restatement must reproduce them
bit for bit -- ids, distances, tree links, Dubins words / lengths / polylines,
collision verdicts.  When oracle/_ref is present (this container) the same
comparison also runs live on fresh random inputs.
"""
import importlib.util
import json
import os

import numpy as np
import pytest

from drrt_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("make_ref_golden", os.path.join(HERE, "golden", "make_ref_golden.py"))
gen = importlib.util.module_from_spec(spec)
spec.loader.exec_module(gen)


def unhex(v):
    return np.array([float.fromhex(x) for x in v], dtype=np.float64)


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(HERE, "golden", "ref_v1.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="module")
def I():
    return gen.inputs()


def test_tree_structure_matches_reference(po, gold, I):
    t = po.RefTree()
    t.insert(I["pos"])
    st = t.structure()
    for k in ("parent", "left", "right", "split"):
        assert st[k].tolist() == gold["structure"][k], k


@pytest.mark.parametrize("name,wrap,metric", [("wrap_r2s1", True, 0), ("nowrap_r2s1", False, 0),
                                              ("wrap_euclid", True, 1)])
def test_range_and_nearest_match_reference(po, gold, I, name, wrap, metric):
    kw = dict(metric=metric) if wrap else dict(wraps=(), wrap_points=(), metric=metric)
    t = po.RefTree(**kw)
    t.insert(I["pos"])
    off, ids, dist = t.range_batch(I["q"], I["rad"])
    g = gold["range_" + name]
    assert off.tolist() == g["off"] and ids.tolist() == g["ids"]
    assert np.array_equal(dist, unhex(g["dist"]))            # JList keys, bit for bit
    nid, nd = t.nearest_batch(I["q"])
    g = gold["nearest_" + name]
    assert nid.tolist() == g["ids"] and np.array_equal(nd, unhex(g["dist"]))


def test_lattice_range_matches_reference(po, gold, I):
    t = po.RefTree(metric=1)
    t.insert(I["grid"])
    off, ids, dist = t.range_batch(I["grid"][::5], 2.0)
    g = gold["range_lattice"]
    assert off.tolist() == g["off"] and ids.tolist() == g["ids"]
    assert np.array_equal(dist, unhex(g["dist"]))


def test_ghosts_match_reference(po, gold, I):
    t = po.RefTree()
    pts = []
    for q, c in zip(I["q"], gold["ghosts"]["counts"]):
        g = t.ghosts(q, 0.6)
        assert len(g) == c
        pts.extend(g.ravel().tolist())
    assert np.array_equal(np.array(pts), unhex(gold["ghosts"]["points"]))


def test_dubins_and_edge_checks_match_reference(po, gold, I):
    g = gold["edges"]
    O = po.RefObstacles(**I["o"])
    v, ln = O.edge_check_batch(I["s"], I["e"])
    assert v.tolist() == g["verdict"]
    assert np.array_equal(ln, unhex(g["dist"]))               # same libm in this container
    for i in range(len(I["s"])):
        tr = po.dubins_trajectory(I["s"][i], I["e"][i])
        assert tr["type"] == g["type"][i] and len(tr["xy"]) == g["npts"][i]
        assert np.array_equal(tr["xy"][:3].ravel(), unhex(g["xy_first"][i]))
        assert np.array_equal(tr["xy"][-2:].ravel(), unhex(g["xy_last"][i]))
    assert "0s0" in g["type"] and 0 < sum(g["verdict"]) < len(g["verdict"])


def test_point_checks_match_reference(po, gold, I):
    O = po.RefObstacles(**I["o"])
    assert O.point_check_batch(I["pts"]).tolist() == gold["points"]


def test_geometry_matches_reference(po, gold):
    seg = unhex(gold["geom"]["segs"]).reshape(-1, 8)
    sd = np.array([po.segment_dist_sqrd(s[0:2], s[2:4], s[4:6], s[6:8]) for s in seg])
    pd = np.array([po.dist_sqrd_point_segment(s[0:2], s[4:6], s[6:8]) for s in seg])
    assert np.array_equal(sd, unhex(gold["geom"]["seg_dist"]))
    assert np.array_equal(pd, unhex(gold["geom"]["pt_dist"]))


def test_live_against_compiled_reference(po):
    """fresh random inputs, only where oracle/_ref was built (needs /root/reference)"""
    from oracle import pyref
    if not pyref.available():
        pytest.skip("oracle/_ref not built here")
    pos = synth.box_points(8000, 0xC1)
    q = synth.box_points(150, 0xC2)
    q[:30, 2] = np.linspace(0, 0.4, 30)
    q[30:60, 2] = synth.TWO_PI - np.linspace(0, 0.4, 30)
    t = po.RefTree()
    t.insert(pos)
    for a, b in zip(pyref.range_batch(pos, q, 1.9), t.range_batch(q, 1.9)):
        assert np.array_equal(a, b)
    for a, b in zip(pyref.nearest_batch(pos, q), t.nearest_batch(q)):
        assert np.array_equal(a, b)
    o = synth.obstacles(48, 0xC3)
    s, e = synth.random_edges(1500, 0xC4, max_len=2.5)
    r = pyref.edges(s, e, o)
    v, ln = po.RefObstacles(**o).edge_check_batch(s, e)
    assert np.array_equal(v, r["verdict"]) and np.array_equal(ln, r["dist"])
    pts = synth.box_points(3000, 0xC5)
    assert np.array_equal(pyref.points(pts, o), po.RefObstacles(**o).point_check_batch(pts))


def test_recreate_command_of_the_reference_driver(po):
    """the create / destroy / create sequence the shim's lifetime test replays
    (tests/test_gpu_shim.py), here over the reference's own kdtree.cpp"""
    from oracle import pyref
    if not pyref.available():
        pytest.skip("oracle/_ref/drrt_ref not built")
    from drrt_b200 import synth
    pos = synth.box_points(600, 0x51)
    q = synth.box_points(40, 0x52)
    a, b = pyref.recreate(pos, q, 4.0)
    for got, m in ((a, 300), (b, 600)):
        t = po.RefTree()
        t.insert(pos[:m])
        roff, rids, _ = t.range_batch(q, 4.0)
        assert got["root"] == 0 and got["size"] == m
        assert np.array_equal(got["off"], roff) and np.array_equal(got["ids"], rids)


def test_timing_commands_of_the_reference_driver(po):
    """`range_time` / `edgecheck_time` (what bench.py --impl reference runs) do the same work as the
    commands checked above: their hit counts equal the oracle's"""
    from oracle import pyref
    if not pyref.available():
        pytest.skip("oracle/_ref not built here")
    pos = synth.box_points(5000, 0xD1)
    q = synth.box_points(40, 0xD2)
    t = po.RefTree()
    t.insert(pos)
    off, _, _ = t.range_batch(q, 2.2)
    hits, secs = pyref.range_time(pos, q, 2.2, warm=1, timed=2)
    assert hits == off[-1] and secs > 0.0
    o = synth.obstacles(32, 0xD3)
    s, e = synth.random_edges(400, 0xD4, max_len=1.0)
    v, _ = po.RefObstacles(**o).edge_check_batch(s, e)
    hits, secs = pyref.edge_time(s, e, o)
    assert hits == int(v.sum()) and secs > 0.0


def _timed_gen():
    spec2 = importlib.util.spec_from_file_location("make_ref_timed_golden",
                                                   os.path.join(HERE, "golden", "make_ref_timed_golden.py"))
    m = importlib.util.module_from_spec(spec2)
    spec2.loader.exec_module(m)
    return m


def test_timed_obstacles_match_reference(po):
    """ExplicitEdgeCheck2D kinds 6 / 7 (obstacle.cpp:805-875): the oracle's restatement against the
    committed verdicts of the compiled reference, and live against the binary when it is here"""
    I = _timed_gen().inputs()
    with open(os.path.join(HERE, "golden", "ref_timed_v1.json")) as fh:
        gold = json.load(fh)
    ro = po.RefTimedObstacles(**I["o"])
    for r in I["radii"]:
        v = ro.edge_check_batch(I["s"], I["e"], r)
        assert v.tolist() == gold["verdict"][str(r)]
        assert 0.02 < v.mean() < 0.6
    from oracle import pyref
    if pyref.available():
        o = synth.timed_obstacles(40, 0xC7)
        s, e = synth.timed_segments(6000, 0xC8)
        assert np.array_equal(pyref.timed(o, s, e, 0.8), po.RefTimedObstacles(**o).edge_check_batch(s, e, 0.8))
