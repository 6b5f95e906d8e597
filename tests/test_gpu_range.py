"""Parity of the CUDA range-ball path (drrt_range_batch, through the C ABI)
against the CPU oracle's restatement of KDTree::KDFindWithinRange
(src/kdtree.cpp:794-820).  Bar: neighbour id sets bit-exact (rows sorted by
id); distances within 1e-5 relative (north star) -- in fact compared exactly.
"""
import numpy as np
import pytest

from drrt_b200 import synth

pytestmark = pytest.mark.gpu
W = synth.TWO_PI


def _check(tree, ref, q, r, exact_dist=True):
    off, ids, dist = tree.KDFindWithinRange(r, q)
    roff, rids, rdist = ref.range_batch(q, r)
    assert np.array_equal(off, roff), "hit counts differ"
    assert np.array_equal(ids, rids), "neighbour ids differ"
    if exact_dist:
        assert np.array_equal(dist, rdist)
    else:
        assert np.allclose(dist, rdist, rtol=1e-5, atol=0)
    return off


def _pair(drrt, po, pos, metric=0, wraps=(2,), wrap_points=(W,)):
    t = drrt.KDTree(wraps=wraps, wrap_points=wrap_points, metric=metric)
    t.KDInsert(pos)
    r = po.RefTree(wraps=wraps, wrap_points=wrap_points, metric=metric)
    r.insert(pos)
    return t, r


def test_range_small_tree_no_index(drrt, po):
    pos = synth.box_points(500, 11)
    t, r = _pair(drrt, po, pos)
    q = synth.box_points(300, 12)
    _check(t, r, q, 6.0)
    _check(t, r, q, 0.5)


def test_range_indexed_sparse_and_dense(drrt, po):
    pos = synth.box_points(200_000, 21)
    t, r = _pair(drrt, po, pos)
    q = synth.box_points(2000, 22)
    off = _check(t, r, q, 1.1)          # warp-per-query mode
    assert off[-1] > 0
    _check(t, r, q[:300], 4.0)          # CTA-per-query mode (~2000 hits each)
    _check(t, r, q[:40], 9.0)           # staging overflow path


def test_range_per_query_radius(drrt, po):
    pos = synth.box_points(100_000, 31)
    t, r = _pair(drrt, po, pos)
    q = synth.box_points(1500, 32)
    rad = np.random.default_rng(33).uniform(0.0, 2.5, len(q))
    rad[:5] = 0.0
    _check(t, r, q, rad)


def test_range_ghost_wraparound(drrt, po):
    # queries hugging theta = 0 and theta = 2*pi' must see neighbours across the seam
    pos = synth.box_points(150_000, 41)
    t, r = _pair(drrt, po, pos)
    q = synth.box_points(3000, 42)
    q[:1000, 2] = np.random.default_rng(43).uniform(0.0, 0.3, 1000)
    q[1000:2000, 2] = W - np.random.default_rng(44).uniform(0.0, 0.3, 1000)
    off = _check(t, r, q, 1.3)
    # the seam matters: dropping the wrap changes the answer
    t2 = drrt.KDTree(wraps=(), wrap_points=())
    t2.KDInsert(pos)
    off2, _, _ = t2.KDFindWithinRange(1.3, q)
    assert off2[-1] <= off[-1]


def test_range_known_answers(drrt, po):
    # SURVEY 8c hand-derived cases
    t = drrt.KDTree()
    t.KDInsert(np.array([[10.0, 10.0, 1.0]]))
    # (1) root at distance exactly r is admitted (kdtree.cpp:802 `<=`)
    off, ids, dist = t.KDFindWithinRange(2.0, np.array([[12.0, 10.0, 1.0]]))
    assert list(ids) == [0] and dist[0] == 2.0
    # the same geometry for a non-root node is a miss (:733 `<`)
    t.KDInsert(np.array([[20.0, 10.0, 1.0]]))
    off, ids, _ = t.KDFindWithinRange(2.0, np.array([[22.0, 10.0, 1.0]]))
    assert len(ids) == 0
    # (2) q theta 0.05 vs node theta 2*pi'-0.05: d = 0.1 through the wrap
    t.KDInsert(np.array([[30.0, 30.0, W - 0.05]]))
    off, ids, dist = t.KDFindWithinRange(0.2, np.array([[30.0, 30.0, 0.05]]))
    assert list(ids) == [2] and abs(dist[0] - 0.1) < 1e-12


def test_range_negative_theta_and_out_of_box(drrt, po):
    # start/goal poses of smalltest.cpp have theta = -3*pi/4 (smalltest.cpp:190-196)
    rng = np.random.default_rng(51)
    pos = synth.box_points(60_000, 52)
    pos[0] = [0.0, 0.0, -3 * synth.PI / 4]
    pos[1] = [49.0, 49.0, -3 * synth.PI / 4]
    pos[100:200, 2] -= W          # negative headings
    pos[200:300, 2] += W          # headings above the wrap point
    pos[300:320, 0] += 80.0       # far outside the bulk
    t, r = _pair(drrt, po, pos)
    q = synth.box_points(1500, 53)
    q[:50] = pos[:50]
    q[50:100, 2] = -rng.uniform(0, 3, 50)
    q[100:150, 2] = W + rng.uniform(0, 3, 50)
    q[150:160, 0] = 120.0
    _check(t, r, q, 1.5)
    _check(t, r, q[:200], 5.0)


def test_range_euclid_metric_thetastar_grid(drrt, po):
    # Theta*'s private tree: 50x50 unit grid, Euclidean metric, theta wraps,
    # range 2 (thetastar.cpp:50-52, 80-110, 152): distances land exactly on r
    xs, ys = np.meshgrid(np.arange(50.0), np.arange(50.0), indexing="ij")
    pos = np.stack([xs.ravel(), ys.ravel(), np.zeros(2500)], 1)
    t, r = _pair(drrt, po, pos, metric=1)
    _check(t, r, pos[::7], 2.0)
    _check(t, r, pos[::7] + [0.25, 0.5, 0.0], 2.0)


def test_range_after_incremental_inserts(drrt, po):
    # the index is rebuilt lazily; nodes in the unsorted tail must still be found
    pos = synth.box_points(30_000, 61)
    t = drrt.KDTree()
    r = po.RefTree()
    q = synth.box_points(64, 62)
    done = 0
    for chunk in (1, 1, 2, 5, 100, 3000, 1, 1, 7, 20_000, 1, 3, 2000, 4878):
        t.KDInsert(pos[done:done + chunk])
        r.insert(pos[done:done + chunk])
        done += chunk
        _check(t, r, q[:8], 3.0)
    assert done == 30_000
    _check(t, r, q, 2.0)


def test_range_empty_inputs(drrt, po):
    t = drrt.KDTree()
    off, ids, dist = t.KDFindWithinRange(1.0, np.zeros((3, 3)))
    assert list(off) == [0, 0, 0, 0] and len(ids) == 0
    t.KDInsert(synth.box_points(10, 1))
    off, ids, dist = t.KDFindWithinRange(1.0, np.zeros((0, 3)))
    assert list(off) == [0]


def test_range_config2_sample(drrt, po):
    # BASELINE config 2 at full node count, on a query sample the oracle finishes in seconds
    pos = synth.nodes()
    t, r = _pair(drrt, po, pos)
    q = synth.queries()[:512]
    for gamma in (synth.GAMMA_SPARSE, synth.GAMMA_REF):
        rad = synth.shrinking_radius(len(pos), gamma)
        _check(t, r, q if gamma == synth.GAMMA_SPARSE else q[:96], rad)


def test_range_clustered_ids_fall_back_to_network_sort(drrt, po):
    # ids of a neighbourhood bunched at both ends of the id range crowd the id
    # buckets of the output sort: the bitonic fallback must give the same rows
    pos = synth.box_points(120_000, 91)
    rng = np.random.default_rng(92)
    blob = lambda n: np.stack([25 + rng.normal(0, 0.2, n), 25 + rng.normal(0, 0.2, n),
                               3 + rng.normal(0, 0.2, n)], 1)
    pos[:120] = blob(120)
    pos[-150:] = blob(150)
    pos[50_000:52_500] = np.stack([10 + rng.normal(0, 0.3, 2500), 10 + rng.normal(0, 0.3, 2500),
                                   2 + rng.normal(0, 0.3, 2500)], 1)
    t, r = _pair(drrt, po, pos)
    q = synth.box_points(400, 93)
    q[:100] = blob(100)
    q[100:200] = [10.0, 10.0, 2.0] + rng.normal(0, 0.2, (100, 3))
    _check(t, r, q, 0.6)     # warp-per-query mode, crowded buckets
    _check(t, r, q, 3.5)     # CTA-per-query mode, crowded buckets


def test_range_cta_crowded_and_wrapped_bucket_counters(drrt, po):
    # CTA-per-query mode sorts by id with BYTE counters per id bucket (id >> shift); ids bunched in one
    # bucket must send the query to the network sort.  With 4.3 M nodes a bucket spans 512 ids: 40
    # consecutive ids in one ball crowd a counter (>= 32), 400 make it wrap (>= 256: the carry lands in
    # the neighbouring counter, and only the counters no longer adding up to the hit count shows it)
    n = 4_300_000
    pos = synth.box_points(n, 94)
    rng = np.random.default_rng(95)
    blob = lambda c, m: np.asarray(c) + rng.normal(0, 0.05, (m, 3))
    pos[2_000_384:2_000_784] = blob([12.0, 30.0, 2.5], 400)      # bucket 3907, ids 2000384 .. 2000895
    pos[3_000_320:3_000_360] = blob([35.0, 14.0, 4.0], 40)       # bucket 5860
    t = drrt.KDTree(capacity=n)
    t.KDInsert(pos)
    r = po.RefTree()
    r.insert(pos)
    q = synth.box_points(24, 96)
    q[:8] = blob([12.0, 30.0, 2.5], 8)
    q[8:16] = blob([35.0, 14.0, 4.0], 8)
    off = _check(t, r, q, 1.45)
    counts = np.diff(off)
    assert counts.min() > 256 and counts.max() <= 4608      # CTA mode, rows staged in shared memory
    assert (counts[:8] > counts[16:].max()).all()           # the blobs are inside their queries' balls


def test_range_signed_prune_without_wrap(drrt, po):
    # a tree created WITHOUT wrapped dimensions still uses the wrap-aware metric;
    # the reference's signed theta prune (kdtree.cpp:744-756) then hides some
    # in-range nodes depending on the tree shape.  The device reproduces that
    # through gate[] (store.cuh), with no ghost pass to repair it.
    pos = synth.box_points(80_000, 95)
    t, r = _pair(drrt, po, pos, wraps=(), wrap_points=())
    q = synth.box_points(3000, 96)
    q[:1500, 2] = W - np.random.default_rng(97).uniform(0.0, 0.5, 1500)
    off = _check(t, r, q, 1.2)
    tw, rw = _pair(drrt, po, pos)
    offw = _check(tw, rw, q, 1.2)
    assert off[-1] < offw[-1]          # the prune really did hide neighbours
    # hand-built case: D sits in the left subtree of the theta-split node C
    p4 = np.array([[10, 10, 3.0], [10, 10, 3.0], [10, 10, 3.1], [10, 10, 0.02]])
    t4, r4 = _pair(drrt, po, p4, wraps=(), wrap_points=())
    _, ids, _ = t4.KDFindWithinRange(0.1, np.array([[10.0, 10.0, W - 0.02]]))
    assert 3 not in ids
    t5, r5 = _pair(drrt, po, p4)
    _, ids, _ = t5.KDFindWithinRange(0.1, np.array([[10.0, 10.0, W - 0.02]]))
    assert list(ids) == [3]


def test_range_signed_prune_follows_later_inserts(drrt, po):
    # the sorted records carry a float copy of gate[]: an insert that gives an INDEXED node its first
    # right theta child lowers that node's gate, and the copy has to follow (no index rebuild in
    # between: the tails stay under the rebuild threshold) -- both kernels, with and without ghost
    pos = synth.box_points(60_000, 98)
    rng = np.random.default_rng(99)
    q = synth.box_points(400, 100)
    q[:300, 2] = W - rng.uniform(0.0, 0.6, 300)
    for wraps, wp in (((), ()), ((2,), (W,))):
        t = drrt.KDTree(wraps=wraps, wrap_points=wp)
        r = po.RefTree(wraps=wraps, wrap_points=wp)
        done = 0
        for chunk in (40_000, 1, 1, 500, 3, 1500, 1, 2000, 90, 1):
            t.KDInsert(pos[done:done + chunk])
            r.insert(pos[done:done + chunk])
            done += chunk
            _check(t, r, q, 1.2)        # warp-per-query kernel
            _check(t, r, q[:48], 6.5)   # CTA-per-query kernel
        t.close()


def _rows_dev(drrt, res):
    """rows of a device-resident result as host arrays: offsets, counts, ids, dist"""
    off, ids, dist = drrt.KDTree.range_result_tensors(res)
    cnt = drrt.KDTree.range_result_counts(res)
    return off.cpu().numpy(), cnt.cpu().numpy(), ids.cpu().numpy(), dist.cpu().numpy()


def test_range_dev_sliced_rows_and_pack(drrt, po):
    # large neighbourhoods are written in slices of consecutive rows (offsets + counts);
    # drrt_range_pack turns the result into the plain CSR the host call returns
    import torch
    pos = synth.box_points(200_000, 71)
    t, r = _pair(drrt, po, pos)
    q = synth.box_points(203, 72)           # not a multiple of the slice length
    roff, rids, rdist = r.range_batch(q, 4.0)
    res = t.KDFindWithinRange_dev(4.0, torch.from_numpy(q).cuda())
    assert res.nq == len(q) and res.total == roff[-1]
    off, cnt, ids, dist = _rows_dev(drrt, res)
    assert np.array_equal(cnt, np.diff(roff))
    assert (off[1:] >= off[:-1] + cnt).all() and off[-1] <= res.extent   # query order, no overlap
    for i in range(len(q)):
        assert np.array_equal(ids[off[i]:off[i] + cnt[i]], rids[roff[i]:roff[i + 1]])
        assert np.array_equal(dist[off[i]:off[i] + cnt[i]], rdist[roff[i]:roff[i + 1]])
    res2 = t.range_pack()
    assert res2.packed == 1 and res2.extent == res2.total == roff[-1]
    off2, cnt2, ids2, dist2 = _rows_dev(drrt, res2)
    assert np.array_equal(off2, roff) and np.array_equal(ids2, rids) and np.array_equal(dist2, rdist)
    # the small-neighbourhood (warp-per-query) kernel writes slices too
    res3 = t.KDFindWithinRange_dev(0.8, torch.from_numpy(q).cuda())
    off3, cnt3, ids3, dist3 = _rows_dev(drrt, res3)
    soff, sids, sdist = r.range_batch(q, 0.8)
    assert np.array_equal(cnt3, np.diff(soff)) and (off3[1:] >= off3[:-1] + cnt3).all()
    for i in range(len(q)):
        assert np.array_equal(ids3[off3[i]:off3[i] + cnt3[i]], sids[soff[i]:soff[i + 1]])
        assert np.array_equal(dist3[off3[i]:off3[i] + cnt3[i]], sdist[soff[i]:soff[i + 1]])
    res4 = t.range_pack()
    off4, _, ids4, dist4 = _rows_dev(drrt, res4)
    assert res4.packed == 1 and np.array_equal(off4, soff) and np.array_equal(ids4, sids) and np.array_equal(dist4, sdist)


def test_range_slice_overflow_reruns_exact(drrt, po):
    # a dense blob makes some slices outgrow their share of the arena: the batch is rerun
    # with exact slice sizes and comes back as a plain CSR
    import torch
    pos = synth.box_points(150_000, 75)
    rng = np.random.default_rng(76)
    pos[20_000:26_000] = [30.0, 20.0, 1.0] + rng.normal(0, 0.25, (6000, 3))
    t, r = _pair(drrt, po, pos)
    q = synth.box_points(120, 77)
    q[40:80] = [30.0, 20.0, 1.0] + rng.normal(0, 0.2, (40, 3))
    roff, rids, rdist = r.range_batch(q, 3.0)
    res = t.KDFindWithinRange_dev(3.0, torch.from_numpy(q).cuda())
    assert res.packed == 1 and res.total == roff[-1]
    off, cnt, ids, dist = _rows_dev(drrt, res)
    assert np.array_equal(off, roff) and np.array_equal(ids, rids) and np.array_equal(dist, rdist)
    _check(t, r, q, 3.0)


def test_range_streamed_one_call(drrt, po):
    """drrt_range_batch_into: the sub-batched, copy-overlapped host call returns exactly
    the two-call result (several sub-batches, per-query radii, CTA mode), reports
    DRRT_ERR_CAPACITY with complete offsets when the rows do not fit, and leaves no
    fetchable result on the handle."""
    from drrt_b200 import capi
    pos = synth.box_points(200_000, 71)
    t, r = _pair(drrt, po, pos)
    q = synth.box_points(40_000, 72)              # three sub-batches of 16384
    rad = np.random.default_rng(73).uniform(0.2, 1.2, len(q))
    for radius in (0.9, rad):
        off, ids, dist = t.KDFindWithinRange(radius, q)
        o2 = np.full(len(q) + 1, -1, dtype=np.int64)
        i2 = np.empty(len(ids) + 7, dtype=np.int32)
        d2 = np.empty(len(ids) + 7, dtype=np.float64)
        total = t.KDFindWithinRangeInto(radius, q, o2, i2, d2)
        assert total == len(ids)
        assert np.array_equal(o2, off)
        assert np.array_equal(i2[:total], ids) and np.array_equal(d2[:total], dist)
    # rows either side of a sub-batch boundary against the oracle itself
    sl = slice(16380, 16390)
    roff, rids, rdist = r.range_batch(q[sl], 0.9)
    i2 = np.empty(1_600_000, dtype=np.int32)
    d2 = np.empty(1_600_000, dtype=np.float64)
    total = t.KDFindWithinRangeInto(0.9, q, o2, i2, d2)
    a, b = o2[sl.start], o2[sl.stop]
    assert np.array_equal(o2[sl.start:sl.stop + 1] - a, roff)
    assert np.array_equal(i2[a:b], rids) and np.array_equal(d2[a:b], rdist)
    # CTA-per-query sub-batches, ids only, pinned torch buffers
    import torch
    qd = synth.box_points(30_000, 74)
    off, ids, dist = t.KDFindWithinRange(4.0, qd)
    o3 = torch.empty(len(qd) + 1, dtype=torch.int64).pin_memory()
    i3 = torch.empty(len(ids), dtype=torch.int32).pin_memory()
    assert t.KDFindWithinRangeInto(4.0, qd, o3, i3, None) == len(ids)
    assert np.array_equal(o3.numpy(), off) and np.array_equal(i3.numpy(), ids)
    # too small: error, offsets complete, a sized retry succeeds
    small = np.empty(len(ids) * 6 // 10, dtype=np.int32)  # the first sub-batch fits, the second does not
    o4 = np.empty(len(qd) + 1, dtype=np.int64)
    with pytest.raises(drrt.capi.DrrtError) as ei:
        t.KDFindWithinRangeInto(4.0, qd, o4, small, None)
    assert ei.value.code == capi.ERR_CAPACITY
    assert np.array_equal(o4, off)
    cut = np.searchsorted(off, len(small), side="right") - 1  # rows of whole sub-batches that fit
    assert cut >= 16384
    full = off[16384]
    assert np.array_equal(small[:full], ids[:full])
    # nothing to fetch afterwards
    lib = capi.load()
    assert lib.drrt_range_fetch(t._h, None, None) == capi.ERR_STATE
    # empty batch
    o5 = np.full(1, -1, dtype=np.int64)
    assert t.KDFindWithinRangeInto(1.0, np.empty((0, 3)), o5, np.empty(0, dtype=np.int32), None) == 0
    assert o5[0] == 0


def test_range_narrow_rows_and_small_batches(drrt, po):
    """drrt_range_batch_into_f32 (ids bit-exact, distances rounded to float: 8 B per hit on the
    wire) and the single-sub-batch path of the host call (the planner's own nq == 1 call,
    src/drrt.cpp:213-216; rows leave straight from the arena when it is already a plain CSR)."""
    from drrt_b200 import capi
    pos = synth.box_points(200_000, 75)
    t, r = _pair(drrt, po, pos)
    q = synth.box_points(40_000, 76)                      # streamed: three sub-batches
    for radius in (0.9, 4.0):
        nq = len(q) if radius < 1 else 20_000
        off, ids, dist = t.KDFindWithinRange(radius, q[:nq])
        o2 = np.full(nq + 1, -1, dtype=np.int64)
        i2 = np.empty(len(ids), dtype=np.int32)
        f2 = np.empty(len(ids), dtype=np.float32)
        assert t.KDFindWithinRangeInto(radius, q[:nq], o2, i2, f2) == len(ids)
        assert np.array_equal(o2, off) and np.array_equal(i2, ids)
        assert np.array_equal(f2, dist.astype(np.float32))          # correctly rounded, not truncated
        assert np.all(np.abs(f2.astype(np.float64) - dist) <= 1e-5 * dist)   # the north star's tolerance
    # small batches: 1 query (one slice: already packed), 300 warp-mode, 300 CTA-mode queries
    for nq, radius in ((1, 0.9), (1, 4.0), (300, 0.9), (300, 4.0), (3, 4.0)):
        roff, rids, rdist = r.range_batch(q[:nq], radius)
        for dtype in (np.float64, np.float32, None):
            o = np.full(nq + 1, -1, dtype=np.int64)
            i = np.empty(len(rids), dtype=np.int32)
            d = None if dtype is None else np.empty(len(rids), dtype=dtype)
            assert t.KDFindWithinRangeInto(radius, q[:nq], o, i, d) == len(rids)
            assert np.array_equal(o, roff) and np.array_equal(i, rids)
            if dtype is not None:
                assert np.array_equal(d, rdist.astype(dtype))
        small = np.empty(max(len(rids) - 1, 0), dtype=np.int32)
        with pytest.raises(drrt.capi.DrrtError) as ei:
            t.KDFindWithinRangeInto(radius, q[:nq], o, small, None)
        assert ei.value.code == capi.ERR_CAPACITY and np.array_equal(o, roff)
    # per-query radii through the small path
    rad = np.linspace(0.3, 4.5, 200)
    roff, rids, rdist = r.range_batch(q[:200], rad)
    o = np.empty(201, dtype=np.int64)
    i = np.empty(len(rids), dtype=np.int32)
    d = np.empty(len(rids), dtype=np.float64)
    assert t.KDFindWithinRangeInto(rad, q[:200], o, i, d) == len(rids)
    assert np.array_equal(o, roff) and np.array_equal(i, rids) and np.array_equal(d, rdist)
