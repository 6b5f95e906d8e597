"""Edge-check parity (drrt_edgecheck_batch through the C ABI) against the
oracle's restatement of CalculateTrajectory + ExplicitEdgeCheck +
ExplicitEdgeCheck2D.  Bar (north star): verdicts identical except edges within
1e-6 of an obstacle boundary; lengths within 1e-9 relative (CUDA vs glibc
atan2/acos differ by a few ulp)."""
import numpy as np
import pytest

from drrt_b200 import synth

pytestmark = pytest.mark.gpu
BAND = 1e-6


def _obs(drrt, po, o):
    t = drrt.KDTree()
    t.SetObstacles(**o)
    ro = po.RefObstacles(**o)
    return t, ro


def _compare(t, ro, s, e, kind=0, radius=0.5, r_min=1.0):
    v, ln = t.ExplicitEdgeCheck(s, e, robot_radius=radius, r_min=r_min, edge_kind=kind)
    rv, rln, clr = ro.edge_check_batch(s, e, edge_kind=kind, robot_radius=radius, r_min=r_min,
                                       want_clearance=True)
    diff = np.nonzero(v != rv)[0]
    assert (clr[diff] < BAND).all(), "verdicts differ outside the 1e-6 boundary band: %s" % diff[:10]
    assert len(diff) <= max(2, len(v) // 100000)
    ok = rln < 1e11
    assert np.allclose(ln[ok], rln[ok], rtol=1e-9, atol=1e-9)
    assert (ln[~ok] > 1e11).all()
    return v, rv


def test_dubins_edges_random(drrt, po):
    o = synth.obstacles()
    t, ro = _obs(drrt, po, o)
    s, e = synth.random_edges(200_000, 81, max_len=0.66)
    v, rv = _compare(t, ro, s, e)
    assert 0.02 < v.mean() < 0.9
    s, e = synth.random_edges(60_000, 82, max_len=5.0)
    _compare(t, ro, s, e)


def test_straight_edges_linecheck(drrt, po):
    o = synth.obstacles()
    t, ro = _obs(drrt, po, o)
    s, e = synth.random_edges(100_000, 83, max_len=3.0)
    _compare(t, ro, s, e, kind=1)


def test_trajectories_match(drrt, po):
    t = drrt.KDTree()
    s, e = synth.random_edges(3000, 84, max_len=3.0)
    xy, npts, names, length = t.CalculateTrajectory(s, e)
    bad = 0
    for i in range(len(s)):
        ref = po.dubins_trajectory(s[i], e[i])
        if names[i] != ref["type"] or npts[i] != len(ref["xy"]):
            bad += 1          # an ulp in atan2 moved ceil(|dphi|/0.1) or a near-tie between words
            continue
        assert np.allclose(xy[i, :npts[i]], ref["xy"], rtol=0, atol=1e-9)
        assert abs(length[i] - ref["dist"]) <= 1e-9 * max(1.0, ref["dist"])
    assert bad <= 3


def test_edge_known_answers(drrt, po):
    # square obstacle [10,20]^2, kind 3
    sq = dict(kind=[3], used=[1], life_span=[1e12], origin=[[15.0, 15.0]], radius=[7.0710678118654755],
              vert_off=[0, 4], verts=[[10, 10], [20, 10], [20, 20], [10, 20]])
    t, ro = _obs(drrt, po, sq)
    th = 0.0
    s = np.array([[14.0, 15.0, th], [5.0, 15.0, th], [0.0, 0.0, th], [9.4, 12.0, 1.5707963]])
    e = np.array([[15.5, 15.0, th], [12.0, 15.0, th], [1.5, 0.0, th], [9.4, 13.5, 1.5707963]])
    # (4) a short straight ("0s0") edge strictly inside the polygon is free (no containment test)
    # crossing the boundary collides; far away is free; passing 0.6 from the wall is free
    v = t.ExplicitEdgeCheck(s, e, want_length=False)
    rv, _ = ro.edge_check_batch(s, e)
    assert list(v) == list(rv)
    assert v[0] == 0 and v[2] == 0
    # (6) unused obstacle / expired obstacle never collide
    for used, life in ((0, 1e12), (1, 0.0)):
        sq2 = dict(sq, used=[used], life_span=[life])
        t.SetObstacles(**sq2)
        assert not t.ExplicitEdgeCheck(s, e, want_length=False).any()
    # (7) kind 1 ball inside the bounding reach collides, kind 2 never
    t.SetObstacles(**dict(sq, kind=[1]))
    v1 = t.ExplicitEdgeCheck(s, e, want_length=False)
    r1, _ = po.RefObstacles(**dict(sq, kind=[1])).edge_check_batch(s, e)
    assert list(v1) == list(r1) and v1[0] == 1
    t.SetObstacles(**dict(sq, kind=[2]))
    assert not t.ExplicitEdgeCheck(s, e, want_length=False).any()
    # (8) in_warmup_time_: every edge is free
    t.SetObstacles(**sq)
    assert not t.ExplicitEdgeCheck(s, e, in_warmup=True, want_length=False).any()
    # kinds 6/7 are refused loudly
    with pytest.raises(drrt.DrrtError):
        t.SetObstacles(**dict(sq, kind=[6]))


def test_edge_ids_variant_matches_pose_variant(drrt, po):
    o = synth.obstacles()
    t, ro = _obs(drrt, po, o)
    pos = synth.box_points(50_000, 85)
    t.KDInsert(pos)
    rng = np.random.default_rng(86)
    a = rng.integers(0, len(pos), 30_000).astype(np.int32)
    off, ids, _ = t.KDFindWithinRange(1.0, pos[a[:2000]])
    b = np.repeat(a[:2000], np.diff(off)).astype(np.int32)
    v1, l1 = t.ExplicitEdgeCheckIds(b, ids)
    v2, l2 = t.ExplicitEdgeCheck(pos[b], pos[ids])
    assert np.array_equal(v1, v2) and np.array_equal(l1, l2)


def test_node_check(drrt, po):
    o = synth.obstacles()
    o["kind"][:20] = 1
    o["kind"][20:30] = 2
    o["kind"][30:40] = 5
    t, ro = _obs(drrt, po, o)
    pts = synth.box_points(200_000, 87)
    v = t.ExplicitNodeCheck(pts, robot_radius=0.5, collision_distance=0.1)
    rv = ro.point_check_batch(pts, robot_radius=0.5, collision_distance=0.1)
    assert np.array_equal(v, rv)
    assert 0.05 < v.mean() < 0.95


def test_edge_round_trip_properties_full_size(drrt, po):
    # size-independent properties on a large batch: no obstacles -> all free;
    # verdicts are independent of batch order; lengths >= straight-line distance
    o = synth.obstacles()
    t, ro = _obs(drrt, po, o)
    s, e = synth.random_edges(2_000_000, 88, max_len=0.66)
    v, ln = t.ExplicitEdgeCheck(s, e)
    perm = np.random.default_rng(89).permutation(len(s))
    v2, ln2 = t.ExplicitEdgeCheck(s[perm], e[perm])
    assert np.array_equal(v[perm], v2) and np.array_equal(ln[perm], ln2)
    d = np.hypot(e[:, 0] - s[:, 0], e[:, 1] - s[:, 1])
    assert (ln >= d * (1 - 1e-12)).all()
    t.SetObstacles(kind=[], used=[], life_span=[], origin=np.zeros((0, 2)), radius=[],
                   vert_off=[0], verts=np.zeros((0, 2)))
    assert not t.ExplicitEdgeCheck(s[:100000], e[:100000], want_length=False).any()


def test_ranked_solve_matches_full_solve(drrt, po):
    # large Dubins batches rank the six candidates in float32 and solve only the ones that can
    # win in double (edgecheck.cu "K4a, ranked"); the lengths must be bit-identical to the
    # solve of all six candidates, on ordinary edges and on near-degenerate ones
    rng = np.random.default_rng(871)
    t = drrt.KDTree()
    t.SetObstacles(**synth.obstacles(32, 872))
    parts = []
    for max_len in (0.3, 0.66, 2.0, 4.1, 9.0):
        parts.append(synth.random_edges(120_000, int(rng.integers(1 << 30)), max_len=max_len))
    s = np.concatenate([p[0] for p in parts])
    e = np.concatenate([p[1] for p in parts])
    k = 40_000
    # existence boundaries and ties: centres 2r / 4r apart, equal headings, mirrored poses
    s[:k, 2] = rng.uniform(0, synth.TWO_PI, k)
    e[:k, 2] = s[:k, 2]                                           # parallel headings: rsr == lsl ties
    d = rng.choice([2.0, 4.0, 1e-7, 1.0], k) + rng.normal(0, 1e-6, k)
    ang = s[:k, 2] + rng.choice([0.0, np.pi / 2, np.pi, -np.pi / 2], k)
    e[:k, 0] = s[:k, 0] + d * np.cos(ang)
    e[:k, 1] = s[:k, 1] + d * np.sin(ang)
    e[k:2 * k, 2] = s[k:2 * k, 2] + np.pi                          # opposite headings
    v, ln = t.ExplicitEdgeCheck(s, e)                              # ranked pipeline (n >= 16384)
    names, full = t.DubinsWords(s, e)                              # all six candidates
    assert np.array_equal(ln, full)
    # and the one-kernel solve used for small batches gives the same verdicts and lengths
    for lo in range(0, 160_000, 16_000):
        v2, ln2 = t.ExplicitEdgeCheck(s[lo:lo + 16_000], e[lo:lo + 16_000])
        assert np.array_equal(ln2, full[lo:lo + 16_000]) and np.array_equal(v2, v[lo:lo + 16_000])
    # verdicts against the oracle on the ordinary edges (the constructed ties above are decided
    # by the last ulp of atan2, where CUDA and glibc may pick different words of equal length)
    ro = po.RefObstacles(**synth.obstacles(32, 872))
    idx = 2 * k + rng.permutation(len(s) - 2 * k)[:60_000]
    rv, rln, clr = ro.edge_check_batch(s[idx], e[idx], want_clearance=True)
    diff = np.nonzero(v[idx] != rv)[0]
    assert (clr[diff] < BAND).all() and len(diff) <= 3
    ok = rln < 1e11
    assert np.allclose(ln[idx][ok], rln[ok], rtol=1e-9, atol=1e-9)


def test_ranked_solve_other_turn_radii(drrt, po):
    # the float32 ranking scales its bands with r_min and the edge length
    t = drrt.KDTree()
    for r_min, max_len, seed in ((0.6, 3.0, 891), (2.5, 1.0, 892), (0.05, 4.0, 893)):
        s, e = synth.random_edges(60_000, seed, max_len=max_len)
        _, ln = t.ExplicitEdgeCheck(s, e, r_min=r_min)          # ranked pipeline
        _, full = t.DubinsWords(s, e, r_min=r_min)              # all six candidates
        assert np.array_equal(ln, full), r_min


def test_long_straight_edges(drrt, po):
    # LineCheck over many grid cells (Theta*'s any-angle edges cross the whole map)
    o = synth.obstacles()
    t, ro = _obs(drrt, po, o)
    rng = np.random.default_rng(895)
    s = synth.box_points(40_000, 896)
    e = synth.box_points(40_000, 897)
    e[:10_000] = s[:10_000] + rng.normal(0, 0.01, (10_000, 3))   # very short ones too
    e[10_000:10_050] = s[10_000:10_050]                          # zero length
    _compare(t, ro, s, e, kind=1)


def test_dubins_edges_poses_at_the_robot_radius(drrt, po):
    """The early decision of the solve kernels (a pose closer than radius - 1e-9 to a listed polygon
    edge decides the edge, edgecheck.cu point_surely_collides): start / end poses placed at
    radius + {0, +-1e-12 ... +-0.2} from polygon edges and vertices -- for polygons whose radius_
    covers them, polygons whose radius_ does not (the bounding reject of obstacle.cpp:765-779 can
    then veto a near edge) and kind-1 balls.  Verdicts may differ from the oracle's only inside the
    1e-6 band; both the ranked pipeline and the small-batch kernel are exercised."""
    o = synth.obstacles(48, 91)
    o["radius"][::3] *= 0.5
    o["kind"][::5] = 1
    t, ro = _obs(drrt, po, o)
    rng = np.random.default_rng(92)
    verts, voff = o["verts"], o["vert_off"]
    offs = np.array([0.0, 1e-12, -1e-12, 1e-10, -1e-10, 1e-8, -1e-8, 2e-6, -2e-6, 1e-3, -1e-3, 0.2, -0.2])
    pts, out = [], []   # points at the radius and the direction pointing away from the boundary there
    for i in range(len(voff) - 1):
        P = verts[voff[i]:voff[i + 1]]
        for k in range(len(P)):
            a, b = P[k - 1], P[k]
            d = b - a
            nrm = np.array([d[1], -d[0]]) / np.hypot(d[0], d[1])
            for tpar in (0.0, 0.3, 1.0):
                for off in offs:
                    for sgn in (1.0, -1.0):
                        pts.append(a + tpar * d + sgn * (0.5 + off) * nrm)
                        out.append(sgn * nrm)
    for i in range(0, len(voff) - 1, 5):   # the balls: reach = robot radius + radius_
        for off in offs:
            ang = rng.uniform(0, 2 * np.pi)
            u = np.array([np.cos(ang), np.sin(ang)])
            pts.append(o["origin"][i] + (0.5 + o["radius"][i] + off) * u)
            out.append(u)
    pts, out = np.array(pts), np.array(out)
    n = len(pts)
    # (a) random headings, partner pose nearby: mostly loops that collide somewhere
    near = np.concatenate([pts, rng.uniform(0, synth.TWO_PI, (n, 1))], 1)
    other = near.copy()
    other[:, :2] += rng.uniform(-0.45, 0.45, (n, 2))
    other[:, 2] = rng.uniform(0, synth.TWO_PI, n)
    # (b) straight runs away from / towards the boundary: the pose at the radius is the only
    # part of the path that can touch it, so the verdict hangs on that pose alone
    ang = np.mod(np.arctan2(out[:, 1], out[:, 0]), synth.TWO_PI)
    far = pts + out * rng.uniform(0.1, 0.6, (n, 1))
    away_s = np.concatenate([pts, ang[:, None]], 1)
    away_e = np.concatenate([far, ang[:, None]], 1)
    back = np.mod(ang + np.pi, synth.TWO_PI)
    to_s = np.concatenate([far, back[:, None]], 1)
    to_e = np.concatenate([pts, back[:, None]], 1)
    s = np.concatenate([near, other, away_s, to_s])
    e = np.concatenate([other, near, away_e, to_e])
    assert len(s) > 16384
    rv, rln, clr = ro.edge_check_batch(s, e, want_clearance=True)
    for lo, hi in ((0, len(s)), (0, 6000)):
        v, ln = t.ExplicitEdgeCheck(s[lo:hi], e[lo:hi])
        diff = lo + np.nonzero(v != rv[lo:hi])[0]
        assert (clr[diff] < BAND).all(), "verdicts differ outside the 1e-6 band: %s" % diff[:10]
    assert 0.2 < rv.mean() < 0.9 and 0.1 < rv[2 * n:].mean() < 0.8
    # poses a full 2e-6 or more outside every boundary are never decided by the shortcut alone:
    # whatever the device answers for them already matched the oracle's walk above


def test_edge_ids_out_of_range_are_refused(drrt):
    """edge end ids are validated on the device, where they are consumed: an id outside the store
    fails the call with ERR_INVALID (small mapped call, one-shot batch and pipelined batch alike)
    instead of reading out of bounds"""
    t = drrt.KDTree()
    t.KDInsert(synth.box_points(5000, 87))
    t.SetObstacles(**synth.obstacles(32, 88))
    for n in (1, 700, 50_000, 2_300_000):
        a = np.random.default_rng(n).integers(0, 5000, n).astype(np.int32)
        b = np.random.default_rng(n + 1).integers(0, 5000, n).astype(np.int32)
        v, ln = t.ExplicitEdgeCheckIds(a, b)                 # all ids valid
        assert len(v) == n
        for bad in (5000, -1):
            a2 = a.copy()
            a2[n // 2] = bad
            with pytest.raises(drrt.DrrtError) as ei:
                t.ExplicitEdgeCheckIds(a2, b)
            assert ei.value.code == drrt.capi.ERR_INVALID
    v2, _ = t.ExplicitEdgeCheckIds(a, b)                      # the handle still works afterwards
    assert np.array_equal(v, v2)


def test_timed_obstacles_kinds_6_7(drrt, po):
    """ExplicitEdgeCheck2D kinds 6 / 7 (obstacle.cpp:805-875) on the device against the oracle and
    the committed verdicts of the compiled reference; verdicts may differ only where the centres'
    closest approach is within 1e-6 of radius_ + radius"""
    import importlib.util
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_ref_timed_golden",
                                                  os.path.join(here, "golden", "make_ref_timed_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    I = gen.inputs()
    with open(os.path.join(here, "golden", "ref_timed_v1.json")) as fh:
        gold = json.load(fh)
    t = drrt.KDTree()
    t.SetTimedObstacles(**I["o"])
    for r in I["radii"]:
        v = t.ExplicitEdgeCheck2DTimed(I["s"], I["e"], r)
        want = np.array(gold["verdict"][str(r)], dtype=np.uint8)
        _, clr = po.RefTimedObstacles(**I["o"]).edge_check_batch(I["s"], I["e"], r, want_clearance=True)
        diff = np.nonzero(v != want)[0]
        assert (clr[diff] < BAND).all() and len(diff) <= 1
    o = synth.timed_obstacles(64, 0xC9)
    s, e = synth.timed_segments(300_000, 0xCA)
    t.SetTimedObstacles(**o)
    v = t.ExplicitEdgeCheck2DTimed(s, e, 0.5)
    rv, clr = po.RefTimedObstacles(**o).edge_check_batch(s, e, 0.5, want_clearance=True)
    diff = np.nonzero(v != rv)[0]
    assert (clr[diff] < BAND).all() and len(diff) <= 3
    assert 0.05 < v.mean() < 0.6
    # the two tables are separate: kinds 6 / 7 are refused by the static table and vice versa
    with pytest.raises(drrt.DrrtError):
        t.SetTimedObstacles(**dict(o, kind=np.full(64, 3)))
    t.SetTimedObstacles(**dict(o, used=np.zeros(64, dtype=np.uint8)))
    assert not t.ExplicitEdgeCheck2DTimed(s[:1000], e[:1000], 0.5).any()


def test_streamed_range_after_pipelined_edges_on_one_handle(drrt, po):
    """The pipelined host edge check and the streamed host range query share the handle's copy
    stream but own their events: either may come first (a range query after a 10 M-edge call once
    found its events missing)."""
    pos = synth.box_points(120_000, 91)
    q = synth.box_points(40_000, 92)                    # > 1.5 sub-batches: the streamed path
    s, e = synth.random_edges(4_300_000, 93)            # > 2 chunks: the pipelined path
    for edges_first in (True, False):
        t = drrt.KDTree()
        t.KDInsert(pos)
        t.SetObstacles(**synth.obstacles(64, 94))
        want = t.KDFindWithinRange(0.9, q)              # two-call form: no streams involved
        for step in ((0, 1) if edges_first else (1, 0)):
            if step == 0:
                v, ln = t.ExplicitEdgeCheck(s, e)
                assert 0.0 < v.mean() < 1.0
            else:
                o = np.empty(len(q) + 1, dtype=np.int64)
                i = np.empty(len(want[1]), dtype=np.int32)
                d = np.empty(len(want[1]), dtype=np.float64)
                assert t.KDFindWithinRangeInto(0.9, q, o, i, d) == len(want[1])
                assert np.array_equal(o, want[0]) and np.array_equal(i, want[1]) and np.array_equal(d, want[2])
        t.close()
