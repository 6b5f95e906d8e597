"""Fused Extend step (drrt_extend_batch) against the oracle:
KDFindWithinRange (kdtree.cpp:794-820), both Dubins edges per neighbour through
CalculateTrajectory + ExplicitEdgeCheck, and FindBestParent's reduction
(drrt.cpp:388-459): the safe new->near edge minimising near.rrt_LMC_ + dist_."""
import numpy as np
import pytest

from drrt_b200 import synth

pytestmark = pytest.mark.gpu


def test_extend_matches_oracle(drrt, po):
    pos = synth.box_points(120_000, 401)
    obstacles = synth.obstacles(128, 402)
    tree = drrt.KDTree()
    tree.KDInsert(pos)
    tree.SetObstacles(**obstacles)
    ref = po.RefTree()
    ref.insert(pos)
    robs = po.RefObstacles(**obstacles)
    rng = np.random.default_rng(403)
    lmc = rng.uniform(0.0, 80.0, len(pos))
    lmc[rng.integers(0, len(pos), 500)] = 1e12              # nodes not yet connected: rrt_LMC_ = INF
    new = synth.box_points(700, 404)
    radius = 1.3
    out = tree.Extend(radius, new, lmc=lmc)
    off, ids, dist = ref.range_batch(new, radius)
    assert np.array_equal(out["offsets"], off) and np.array_equal(out["ids"], ids)
    assert np.array_equal(out["dist"], dist)
    rows = np.repeat(np.arange(len(new)), np.diff(off))
    a, b = new[rows], pos[ids]
    v_out, l_out, c_out = robs.edge_check_batch(a, b, want_clearance=True)   # new -> near
    v_in, l_in, c_in = robs.edge_check_batch(b, a, want_clearance=True)      # near -> new
    for col, (v, ln, clr) in enumerate(((v_out, l_out, c_out), (v_in, l_in, c_in))):
        bad = np.nonzero(out["unsafe"][:, col] != v)[0]
        assert (clr[bad] < 1e-6).all() and len(bad) <= 3
        assert np.allclose(out["length"][:, col], ln, rtol=1e-9)
    # FindBestParent on the DEVICE's own verdicts/lengths must be the exact argmin;
    # against the oracle's it may differ only where a verdict differed
    cost_dev = np.where(out["unsafe"][:, 0] == 0, lmc[ids] + out["length"][:, 0], np.inf)
    for i in range(len(new)):
        sl = slice(off[i], off[i + 1])
        c = cost_dev[sl]
        ok = c < 1e12                                          # rrt_LMC_ starts at INF, strict `>`
        if ok.any():
            j = int(np.argmin(np.where(ok, c, np.inf)))
            assert out["best_parent"][i] == ids[sl][j]
            assert out["best_cost"][i] == c[j]
        else:
            assert out["best_parent"][i] == -1 and out["best_cost"][i] == 1e12
    cost_ref = np.where(v_out == 0, lmc[ids] + l_out, np.inf)
    agree = 0
    for i in range(len(new)):
        sl = slice(off[i], off[i + 1])
        c = cost_ref[sl]
        want = ids[sl][int(np.argmin(c))] if (c < 1e12).any() else -1
        agree += int(want == out["best_parent"][i])
    assert agree >= len(new) - 3
    assert (out["best_parent"] >= 0).mean() > 0.5


def test_extend_warmup_and_empty(drrt, po):
    pos = synth.box_points(5000, 405)
    tree = drrt.KDTree()
    tree.KDInsert(pos)
    tree.SetObstacles(**synth.obstacles(64, 406))
    new = synth.box_points(50, 407)
    out = tree.Extend(2.0, new, lmc=np.zeros(len(pos)), in_warmup=True)
    assert not out["unsafe"].any()                            # in_warmup_time_: nothing collides
    out = tree.Extend(1e-9, new, lmc=np.zeros(len(pos)))
    assert out["offsets"][-1] == 0 and (out["best_parent"] == -1).all()
    out = tree.Extend(2.0, new)                               # no lmc: rows only
    assert out["best_parent"] is None and len(out["ids"]) == out["offsets"][-1]


def test_extend_against_the_resident_lmc_mirror(drrt, po):
    """drrt_lmc_write / drrt_lmc_scatter keep a device mirror of rrt_LMC_; drrt_extend_batch with
    lmc == NULL reduces against it (nothing store-sized crosses PCIe per call) and returns what
    a whole-array upload returns.  Nodes never written read INF (drrt.cpp:391)."""
    from drrt_b200 import capi
    pos = synth.box_points(90_000, 411)
    tree = drrt.KDTree()
    tree.KDInsert(pos[:60_000])
    tree.SetObstacles(**synth.obstacles(96, 412))
    rng = np.random.default_rng(413)
    lmc = rng.uniform(0.0, 60.0, 60_000)
    new = synth.box_points(500, 414)
    want = tree.Extend(1.2, new, lmc=lmc)
    # (a) mirror written in two ranges
    tree.WriteLMC(0, lmc[:25_000])
    tree.WriteLMC(25_000, lmc[25_000:])
    assert np.array_equal(tree.ReadLMC(), lmc)
    cnt, best, cost, total = tree.ExtendParents(1.2, new)
    assert np.array_equal(cnt, np.diff(want["offsets"])) and total == len(want["ids"])
    assert np.array_equal(best, want["best_parent"]) and np.array_equal(cost, want["best_cost"])
    # (b) a rewiring cascade changes some entries: delta update only
    ch = rng.choice(60_000, 4000, replace=False).astype(np.int32)
    lmc[ch] = rng.uniform(0.0, 10.0, len(ch))
    tree.ScatterLMC(ch, lmc[ch])
    want = tree.Extend(1.2, new, lmc=lmc)
    cnt, best, cost, _ = tree.ExtendParents(1.2, new)
    assert np.array_equal(best, want["best_parent"]) and np.array_equal(cost, want["best_cost"])
    # (c) the store grows: new nodes read INF until written, so they are never chosen
    tree.KDInsert(pos[60_000:])
    full = np.concatenate([lmc, np.full(30_000, 1e12)])
    assert np.array_equal(tree.ReadLMC(), full)
    want = tree.Extend(1.2, new, lmc=full)
    cnt, best, cost, _ = tree.ExtendParents(1.2, new)
    assert np.array_equal(best, want["best_parent"]) and np.array_equal(cost, want["best_cost"])
    assert (best < 60_000).all()
    with pytest.raises(drrt.DrrtError):
        tree.ScatterLMC(np.array([90_000], dtype=np.int32), [1.0])
    with pytest.raises(ValueError):
        tree.Extend(1.2, new, lmc=lmc)            # a short host array is refused, not read past its end
    # a later range call replaces the rows: a late fetch is refused, never served from freed memory
    tree.Extend(1.2, new, lmc=full)
    tree.KDFindWithinRange(3.0, new)
    lib = capi.load()
    assert lib.drrt_extend_fetch(tree._h, None, None, None, None) == capi.ERR_STATE


def test_extend_rewire_candidates(drrt):
    """drrt_extend_rewire: Extend's rewiring test (drrt.cpp:330-333) evaluated on the device over
    the pairs of the last fused Extend; only the candidates cross PCIe.  Against the same predicate
    in numpy over the rows / verdicts / lengths the Extend returned."""
    pos = synth.box_points(60_000, 421)
    tree = drrt.KDTree()
    tree.KDInsert(pos)
    tree.SetObstacles(**synth.obstacles(64, 422))
    rng = np.random.default_rng(423)
    lmc = rng.uniform(0.0, 70.0, len(pos))
    lmc[rng.integers(0, len(pos), 300)] = 1e12
    new = synth.box_points(400, 424)
    for goal_lmc in (1e12, 45.0):
        out = tree.Extend(1.4, new, lmc=lmc)
        qi, ni, lm = tree.ExtendRewire(goal_lmc)
        off, ids = out["offsets"], out["ids"]
        row = np.repeat(np.arange(len(new)), np.diff(off))
        via = out["best_cost"][row] + out["length"][:, 1]
        want = (out["unsafe"][:, 1] == 0) & (lmc[ids] > via) & (out["best_parent"][row] != ids) & (via < goal_lmc)
        t = np.nonzero(want)[0]
        assert np.array_equal(qi, row[t]) and np.array_equal(ni, ids[t]) and np.array_equal(lm, via[t])
        assert 0 < len(t) < len(ids)
    # the mirror path gives the same candidates
    tree.WriteLMC(0, lmc)
    tree.ExtendParents(1.4, new)
    q2, n2, l2 = tree.ExtendRewire(45.0)
    assert np.array_equal(q2, qi) and np.array_equal(n2, ni) and np.array_equal(l2, lm)
    # without a parent reduction there is nothing to rewire from
    tree.Extend(1.4, new)
    with pytest.raises(drrt.DrrtError):
        tree.ExtendRewire(45.0)
