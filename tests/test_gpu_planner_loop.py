"""BASELINE config 1 (plumbing): the reference's planner iteration driven through
the C ABI one call at a time, with the parameter block of src/smalltest.cpp:180-220
(box [0,50]^2 x [0,2pi'], start (0,0,-3pi/4), goal (49,49,-3pi/4), delta 5,
gamma 100, robot radius 0.5, r_min 1.0, theta wraps at 2pi') and a fixed RNG (the
reference's sampler is unseedable).  Every iteration does what
RrtMainLoop + Extend do on the path (mainloop.cpp:204-256, drrt.cpp:213-293):

    sample -> KDFindNearest -> saturate -> ExplicitNodeCheck -> KDFindWithinRange
           -> 2*|N| Dubins edge checks -> KDInsert

on the device AND on the CPU oracle, comparing every answer in the loop."""
import math

import numpy as np
import pytest

from drrt_b200 import synth

pytestmark = pytest.mark.gpu
W = synth.TWO_PI


def saturate(new, near, delta, dist):
    # Edge::Saturate dubinsedge.cpp:24-35
    out = new.copy()
    out[:2] = near[:2] + (new[:2] - near[:2]) * delta / dist
    while out[2] < -W:
        out[2] += W
    while out[2] > W:
        out[2] -= W
    return out


def test_planner_loop_in_step_with_oracle(drrt, po):
    iters = 10_000     # >= 10 K inserts
    delta, gamma = synth.DELTA, synth.GAMMA_REF
    rng = np.random.default_rng(0xD2270011)
    obstacles = synth.obstacles(24, 0xD2270012, lo=8.0, hi=42.0)
    tree = drrt.KDTree()
    ref = po.RefTree()
    tree.SetObstacles(**obstacles)
    robs = po.RefObstacles(**obstacles)
    root = np.array([[0.0, 0.0, -3 * synth.PI / 4]])       # smalltest.cpp:190-196 (tree grows from the goal side)
    tree.KDInsert(root)
    ref.insert(root)
    n_edges = n_hits = n_coll = 0
    for it in range(iters):
        n = tree.tree_size_
        radius = synth.shrinking_radius(n, gamma, delta)
        sample = np.array([rng.uniform(0, 50), rng.uniform(0, 50), rng.uniform(0, W)])
        nid, nd = tree.KDFindNearest(sample[None])
        rid, rd = ref.nearest_batch(sample[None], nthreads=1)
        assert nid[0] == rid[0] and abs(nd[0] - rd[0]) <= 1e-5 * rd[0]
        near = tree.positions(int(nid[0]), 1)[0]
        new = saturate(sample, near, delta, nd[0]) if nd[0] > delta else sample
        blocked = tree.ExplicitNodeCheck(new[None])
        assert blocked[0] == robs.point_check_batch(new[None], nthreads=1)[0]
        if blocked[0]:
            continue
        off, ids, dist = tree.KDFindWithinRange(radius, new[None])
        roff, rids, rdist = ref.range_batch(new[None], radius, nthreads=1)
        assert np.array_equal(ids, rids) and np.array_equal(dist, rdist)
        n_hits += len(ids)
        if len(ids):
            nbr = ref_positions(ref, po)[ids]
            s = np.concatenate([np.repeat(new[None], len(ids), 0), nbr])
            e = np.concatenate([nbr, np.repeat(new[None], len(ids), 0)])
            v, ln = tree.ExplicitEdgeCheck(s, e)
            rv, rln, clr = robs.edge_check_batch(s, e, want_clearance=True, nthreads=1)
            bad = np.nonzero(v != rv)[0]
            assert (clr[bad] < 1e-6).all()
            ok = rln < 1e11
            assert np.allclose(ln[ok], rln[ok], rtol=1e-9)
            n_edges += len(v)
            n_coll += int(v.sum())
        assert tree.KDInsert(new[None]) == ref.insert(new[None]) == n
    assert tree.tree_size_ > iters // 2 and n_edges > 10_000 and 0 < n_coll < n_edges
    a, b = tree.structure(), ref.structure()
    for k in a:
        assert np.array_equal(a[k], b[k])


def ref_positions(ref, po):
    import ctypes as C
    n = ref.size
    po.lib().ref_tree_positions.restype = C.POINTER(C.c_double)
    po.lib().ref_tree_positions.argtypes = [C.c_void_p]
    p = po.lib().ref_tree_positions(ref._h)
    return np.ctypeslib.as_array(p, shape=(n, 3))
