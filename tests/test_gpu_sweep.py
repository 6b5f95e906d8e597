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
