"""nearest / k-nearest parity (drrt_nearest_batch, drrt_knn_batch) against the
oracle: KDTree::KDFindNearest src/kdtree.cpp:264-307 restated as a tree walk,
and k-nearest by intent as brute force."""
import numpy as np
import pytest

from drrt_b200 import synth

pytestmark = pytest.mark.gpu


def _pair(drrt, po, pos, **kw):
    t = drrt.KDTree(**kw); t.KDInsert(pos)
    r = po.RefTree(**kw); r.insert(pos)
    return t, r


def test_nearest_matches_tree_walk(drrt, po):
    pos = synth.box_points(200_000, 71)
    t, r = _pair(drrt, po, pos)
    q = synth.box_points(4000, 72)
    q[:500, 2] = np.random.default_rng(1).uniform(0, 0.05, 500)       # seam
    q[500:1000, 2] = synth.TWO_PI - np.random.default_rng(2).uniform(0, 0.05, 500)
    ids, dist = t.KDFindNearest(q)
    rids, rdist = r.nearest_batch(q)
    assert np.array_equal(ids, rids)
    assert np.allclose(dist, rdist, rtol=1e-5, atol=0)


def test_knn_matches_brute_force(drrt, po):
    pos = synth.box_points(20_000, 73)
    t, r = _pair(drrt, po, pos)
    q = synth.box_points(300, 74)
    for k in (1, 4, 16, 33, 100):
        ids, dist = t.KDFindKNearest(k, q)
        rids, rdist = r.knn_batch(q, k)
        assert np.array_equal(ids, rids), k
        assert np.array_equal(dist, rdist), k


def test_knn_ties_break_by_id(drrt, po):
    # duplicate poses: equal distances must come back in id order
    base = synth.box_points(500, 75)
    pos = np.concatenate([base, base, base])
    t, r = _pair(drrt, po, pos)
    ids, dist = t.KDFindKNearest(6, base[:50])
    rids, rdist = r.knn_batch(base[:50], 6)
    assert np.array_equal(ids, rids)
    assert (ids[:, 0] < ids[:, 1]).all() and (dist[:, 0] == 0).all()


def test_knn_fewer_nodes_than_k(drrt, po):
    pos = synth.box_points(5, 76)
    t, r = _pair(drrt, po, pos)
    ids, dist = t.KDFindKNearest(8, synth.box_points(3, 77))
    rids, rdist = r.knn_batch(synth.box_points(3, 77), 8)
    assert np.array_equal(ids, rids)
    assert (ids[:, 5:] == -1).all()


def test_knn_full_scale_sample(drrt, po):
    pos = synth.nodes()
    t = drrt.KDTree(); t.KDInsert(pos)
    r = po.RefTree(); r.insert(pos)
    q = synth.queries()[:64]
    ids, dist = t.KDFindKNearest(16, q)
    rids, rdist = r.knn_batch(q, 16)
    assert np.array_equal(ids, rids) and np.array_equal(dist, rdist)
    nid, nd = t.KDFindNearest(synth.queries()[:2000])
    # property at full size: nearest == first of k-nearest, and it is within its own range ball
    assert np.array_equal(nid[:64], ids[:, 0])
    off, rid, rd = t.KDFindWithinRange(nd * (1 + 1e-12) + 1e-12, synth.queries()[:2000])
    assert (np.diff(off) >= 1).all()


def test_knn_crowded_neighbourhoods_use_the_insertion_kernel(drrt, po):
    # the collect-then-select kernel stages at most 256 nodes per query; queries inside a dense
    # blob exceed that at the first radius and are answered by the insertion kernel instead
    pos = synth.box_points(40_000, 78)
    rng = np.random.default_rng(79)
    pos[10_000:16_000] = [25.0, 25.0, 3.0] + rng.normal(0, 0.15, (6000, 3))
    t, r = _pair(drrt, po, pos)
    q = synth.box_points(400, 80)
    q[:200] = [25.0, 25.0, 3.0] + rng.normal(0, 0.1, (200, 3))
    for k in (1, 16, 64):
        ids, dist = t.KDFindKNearest(k, q)
        rids, rdist = r.knn_batch(q, k)
        assert np.array_equal(ids, rids), k
        assert np.array_equal(dist, rdist), k
