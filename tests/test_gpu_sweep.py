"""Obstacle add / remove sweeps through the C ABI against the oracle:
FindPointsInConflictWithObstacle (src/obstacle.cpp:540-610, Dubins case :554-559), the
single-obstacle re-check `edge->ExplicitEdgeCheck(O)` of AddObstacle (:612-667) and
RemoveObstacle (:669-754), and RemoveObstacle's "any other live obstacle" re-check (:717-734).
The oracle checks a subset by clearing obstacle_used_ of the others: ExplicitEdgeCheck2D
answers false for those (:761), which is what skipping them in the loop does."""
import numpy as np
import pytest

from drrt_b200 import synth

pytestmark = pytest.mark.gpu
BAND = 1e-6
PI = synth.PI


def _subset(po, o, mask):
    m = dict(o)
    m["used"] = (np.asarray(o["used"]) != 0) & (np.asarray(mask) != 0)
    return po.RefObstacles(**m)


def _agree(v, rv, clr):
    diff = np.nonzero(v != rv)[0]
    assert (clr[diff] < BAND).all(), "verdicts differ outside the boundary band: %s" % diff[:10]
    assert len(diff) <= 2


def test_conflict_queries_match_oracle(drrt, po):
    pos = synth.box_points(150_000, 501)
    o = synth.obstacles(64, 502)
    t = drrt.KDTree()
    t.KDInsert(pos)
    t.SetObstacles(**o)
    r = po.RefTree()
    r.insert(pos)
    which = np.array([0, 5, 17, 63, 5], dtype=np.int32)
    off, ids, dist = t.FindPointsInConflictWithObstacle(which, robot_radius=0.5)
    centres = np.concatenate([o["origin"][which], np.full((len(which), 1), PI)], 1)
    radii = 0.5 + o["radius"][which] + PI
    roff, rids, rdist = r.range_batch(centres, radii)
    assert np.array_equal(off, roff) and np.array_equal(ids, rids) and np.array_equal(dist, rdist)
    assert off[-1] > 1000


def test_single_obstacle_recheck_add_obstacle(drrt, po):
    # AddObstacle: the out-edges of the nodes in conflict are tested against O alone
    pos = synth.box_points(80_000, 511)
    o = synth.obstacles(96, 512)
    o["kind"] = np.array(o["kind"]).copy()
    o["kind"][10:14] = 1                       # a few balls take the other branch (:781)
    t = drrt.KDTree()
    t.KDInsert(pos)
    t.SetObstacles(**o)
    rng = np.random.default_rng(513)
    for ob in (3, 11, 40):
        off, ids, _ = t.FindPointsInConflictWithObstacle([ob])
        near = ids[rng.permutation(len(ids))[:3000]]
        # neighbour edges: to a node within a metre or so
        other = near[rng.permutation(len(near))]
        mask = np.zeros(96, dtype=np.uint8)
        mask[ob] = 1
        v, ln = t.ExplicitEdgeCheckSubset(mask, start_ids=near, end_ids=other)
        rv, rln, clr = _subset(po, o, mask).edge_check_batch(pos[near], pos[other], want_clearance=True)
        _agree(v, rv, clr)
        v2, _ = t.ExplicitEdgeCheckSubset(mask, start=pos[near], end=pos[other])
        assert np.array_equal(v, v2)
        ok = rln < 1e11
        assert np.allclose(ln[ok], rln[ok], rtol=1e-9, atol=1e-9)
        # the full check can only add hits
        vall, _ = t.ExplicitEdgeCheckIds(near, other)
        assert (vall >= v).all()


def test_other_obstacles_recheck_remove_obstacle(drrt, po):
    # RemoveObstacle: edges that hit O are tested against every other used obstacle whose
    # life window contains time_elapsed (the host builds that mask)
    o = synth.obstacles(128, 521)
    o["used"] = np.array(o["used"]).copy()
    o["used"][::7] = 0
    t = drrt.KDTree()
    t.SetObstacles(**o)
    s, e = synth.random_edges(60_000, 522, max_len=2.0)
    rng = np.random.default_rng(523)
    in_window = rng.random(128) < 0.8
    for ob in (1, 64):
        mask = (np.asarray(o["used"]) != 0) & in_window
        mask[ob] = False
        v, _ = t.ExplicitEdgeCheckSubset(mask.astype(np.uint8), start=s, end=e)
        rv, _, clr = _subset(po, o, mask).edge_check_batch(s, e, want_clearance=True)
        _agree(v, rv, clr)
        assert 0.01 < v.mean() < 0.99
    # straight edges and the empty subset
    v, _ = t.ExplicitEdgeCheckSubset(np.zeros(128, dtype=np.uint8), start=s[:5000], end=e[:5000])
    assert not v.any()
    mask = np.zeros(128, dtype=np.uint8)
    mask[5:9] = 1
    v, _ = t.ExplicitEdgeCheckSubset(mask, start=s[:20000], end=e[:20000], edge_kind=drrt.EDGE_STRAIGHT)
    rv, _, clr = _subset(po, o, mask).edge_check_batch(s[:20000], e[:20000], edge_kind=1, want_clearance=True)
    _agree(v, rv, clr)


def test_subset_mask_must_match_table(drrt, po):
    t = drrt.KDTree()
    t.SetObstacles(**synth.obstacles(16, 531))
    s, e = synth.random_edges(10, 532)
    with pytest.raises(drrt.capi.DrrtError):
        t.ExplicitEdgeCheckSubset(np.ones(15, dtype=np.uint8), start=s, end=e)
    with pytest.raises(drrt.capi.DrrtError):
        t.FindPointsInConflictWithObstacle([16])


def _stored_edges(t, pos, radius, n_src, seed):
    """the planner's neighbour lists as a stored-edge list: every node of a sample links to its
    neighbours within `radius` (drrt.cpp:256-354), both directions"""
    rng = np.random.default_rng(seed)
    src = np.sort(rng.choice(len(pos), n_src, replace=False)).astype(np.int32)
    off, ids, _ = t.KDFindWithinRange(radius, pos[src])
    a = np.repeat(src, np.diff(off)).astype(np.int32)
    b = ids.astype(np.int32)
    keep = a != b
    a, b = a[keep], b[keep]
    return np.concatenate([a, b]), np.concatenate([b, a])     # out-edges, then the reverse edges


def test_device_resident_edge_list_sweeps(drrt, po):
    """AddObstacle / RemoveObstacle (obstacle.cpp:612-754) in one call each over the device-resident
    edge list: conflict query, selection of the stored edges starting in a conflict node,
    edge->ExplicitEdgeCheck(O) and the other-obstacles re-check, against the oracle doing the same
    steps on the host."""
    pos = synth.box_points(120_000, 541)
    o = synth.obstacles(96, 542)
    o["kind"] = np.array(o["kind"]).copy()
    o["kind"][20:24] = 1
    o["used"] = np.array(o["used"]).copy()
    o["used"][::9] = 0
    t = drrt.KDTree()
    t.KDInsert(pos)
    t.SetObstacles(**o)
    r = po.RefTree()
    r.insert(pos)
    eu, ev = _stored_edges(t, pos, 0.9, 30_000, 543)
    half = len(eu) // 3
    assert t.AppendEdges(eu[:half], ev[:half]) == 0           # appended as the tree grows
    assert t.AppendEdges(eu[half:], ev[half:]) == half
    assert t.stored_edges_ == len(eu) > 200_000
    in_window = np.random.default_rng(544).random(96) < 0.85
    for ob in (2, 22, 50):
        # the oracle's version of the sweep
        centre = np.array([[o["origin"][ob][0], o["origin"][ob][1], PI]])
        _, conflict, _ = r.range_batch(centre, 0.5 + o["radius"][ob] + PI)
        marked = np.zeros(len(pos), dtype=bool)
        marked[conflict] = True
        sel = np.nonzero(marked[eu])[0]
        only = np.zeros(96, dtype=np.uint8)
        only[ob] = 1
        rv1, _, clr1 = _subset(po, o, only).edge_check_batch(pos[eu[sel]], pos[ev[sel]], want_clearance=True)
        others = ((np.asarray(o["used"]) != 0) & in_window).astype(np.uint8)
        others[ob] = 0
        rv2, _, clr2 = _subset(po, o, others).edge_check_batch(pos[eu[sel]], pos[ev[sel]], want_clearance=True)
        # AddObstacle
        add = t.ObstacleSweep(ob)
        assert add["conflict_nodes"] == len(conflict) and add["edges_checked"] == len(sel)
        got = np.zeros(len(eu), dtype=np.uint8)
        got[add["edge_index"]] = 1
        assert (np.diff(add["edge_index"]) > 0).all() and (add["flags"] == 1).all()
        _agree(got[sel], rv1, clr1)
        assert not got[~marked[eu]].any()                     # only edges starting in a conflict node
        # RemoveObstacle: the same edges, flagged when another live obstacle still blocks them
        mask_in = others.copy()
        mask_in[ob] = 1                                        # the entry of O itself is ignored
        rem = t.ObstacleSweep(ob, other_mask=mask_in)
        assert np.array_equal(rem["edge_index"], add["edge_index"])
        blocked = np.zeros(len(eu), dtype=np.uint8)
        blocked[rem["edge_index"]] = (rem["flags"] >> 1) & 1
        hit = got[sel] == 1
        _agree(blocked[sel][hit], rv2[hit], clr2[hit])
        if o["used"][ob]:
            assert 0 < len(add["edge_index"]) < len(sel)
        else:
            assert len(add["edge_index"]) == 0                # an unused obstacle never collides (:761)
    # straight stored edges (Theta*'s LineCheck graph) take the same path
    add = t.ObstacleSweep(50, edge_kind=drrt.EDGE_STRAIGHT)
    assert add["edges_checked"] > 0
    t.ClearEdges()
    assert t.stored_edges_ == 0 and t.ObstacleSweep(2)["edges_checked"] == 0
    with pytest.raises(drrt.DrrtError):
        t.AppendEdges([0], [len(pos)])                        # an end outside the store


def test_moving_obstacles_update_the_device_table(drrt, po):
    """Obstacle::UpdateObstacles -> NextOrigin (obstacle.cpp:12-22, 182-195): origin_ of a moving
    obstacle jumps along its path.  drrt_obstacles_move patches the listed entries; the verdicts
    then equal a fresh table's -- both when the origin alone moves (what the analytic checker sees,
    SURVEY F10) and when the polygon moves with it -- and drrt_obstacles_update switches obstacles
    off and on as the obstacle thread does."""
    o = synth.obstacles(64, 551)
    t = drrt.KDTree()
    t.SetObstacles(**o)
    s, e = synth.random_edges(60_000, 552, max_len=1.5)
    pts = synth.box_points(20_000, 553)
    rng = np.random.default_rng(554)
    for step in range(3):
        idx = rng.choice(64, 9, replace=False).astype(np.int32)
        shift = rng.choice([-1.0, 1.0], (9, 2))
        translate = step != 1
        o = {k: np.array(v).copy() for k, v in o.items()}
        if translate:
            for i, d in zip(idx, shift):
                o["verts"][o["vert_off"][i]:o["vert_off"][i + 1]] += d
        o["origin"][idx] += shift
        t.MoveObstacles(idx, o["origin"][idx], translate_polygon=translate)
        ro = po.RefObstacles(**o)
        v, ln = t.ExplicitEdgeCheck(s, e)
        rv, rln, clr = ro.edge_check_batch(s, e, want_clearance=True)
        _agree(v, rv, clr)
        fresh = drrt.KDTree()
        fresh.SetObstacles(**o)
        assert np.array_equal(v, fresh.ExplicitEdgeCheck(s, e)[0])
        assert np.array_equal(t.ExplicitNodeCheck(pts), ro.point_check_batch(pts))
        assert np.array_equal(t.LineCheck(s[:20_000], e[:20_000]), fresh.LineCheck(s[:20_000], e[:20_000]))
    # obstacle_used_ / life_span_ of a few entries
    off_idx = np.array([3, 30, 31], dtype=np.int32)
    t.UpdateObstacles(off_idx, used=[0, 1, 1], life_span=[1e12, 0.0, 1e12])
    o["used"][3] = 0
    o["life_span"][30] = 0.0
    rv, _, clr = po.RefObstacles(**o).edge_check_batch(s, e, want_clearance=True)
    _agree(t.ExplicitEdgeCheck(s, e)[0], rv, clr)
    with pytest.raises(drrt.DrrtError):
        t.MoveObstacles([64], [[0.0, 0.0]])
