"""A small pass over every kernel of the library, sized for compute-sanitizer:
    compute-sanitizer --tool memcheck  python benchmarks/sanitizer_pass.py
    compute-sanitizer --tool racecheck python benchmarks/sanitizer_pass.py
(insert, both range kernels, the streamed one-call range query over several sub-batches, nearest /
k-nearest, ranked and small-batch Dubins checks, LineCheck, the obstacle-subset check, fused Extend; narrow rows,
the mapped single-call paths, the rrt_LMC_ mirror and rewiring test, stored-edge sweeps, obstacle moves, timed obstacles)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drrt_b200  # noqa: E402
from drrt_b200 import synth  # noqa: E402

scale = int(os.environ.get("SANITIZER_SCALE", "1"))
t = drrt_b200.KDTree(device=0)
t.KDInsert(synth.box_points(30000 * scale, 1))
t.KDInsert(synth.box_points(700, 2))           # an unsorted tail beside the index
q = synth.box_points(26000, 3)
off, ids, dist = t.KDFindWithinRange(0.8, q)            # warp kernel
off2 = np.empty(len(q) + 1, dtype=np.int64)
ids2 = np.empty(len(ids), dtype=np.int32)
dist2 = np.empty(len(ids), dtype=np.float64)
assert t.KDFindWithinRangeInto(0.8, q, off2, ids2, dist2) == len(ids)   # two sub-batches
assert np.array_equal(off, off2) and np.array_equal(ids, ids2) and np.array_equal(dist, dist2)
offc, idsc, _ = t.KDFindWithinRange(6.0, q[:300])       # CTA kernel
t.KDFindNearest(q[:2000])
t.KDFindKNearest(8, q[:2000])
t.SetObstacles(**synth.obstacles(64, 4))
s, e = synth.random_edges(20000, 5, max_len=0.66)
v, ln = t.ExplicitEdgeCheck(s, e)                       # ranked pipeline + work list
v2, _ = t.ExplicitEdgeCheck(s[:3000], e[:3000])         # one-kernel solve
assert np.array_equal(v[:3000], v2)
t.ExplicitEdgeCheck(s[:5000], e[:5000], edge_kind=drrt_b200.EDGE_STRAIGHT)
mask = np.zeros(64, dtype=np.uint8)
mask[::2] = 1
t.ExplicitEdgeCheckSubset(mask, start=s[:17000], end=e[:17000])
t.Extend(0.7, synth.box_points(300, 6))
t.ExplicitNodeCheck(q[:5000])
# ---- round 2: narrow rows, mapped single-call paths, lmc mirror + rewiring test, stored-edge sweeps,
# obstacle moves, time-parametrised obstacles, id validation
f32 = np.empty(len(ids), dtype=np.float32)
assert t.KDFindWithinRangeInto(0.8, q, off2, ids2, f32) == len(ids)
assert t.KDFindWithinRangeInto(0.8, q, off2, ids2, None) == len(ids)
for k in range(20):
    o1 = np.empty(2, dtype=np.int64)
    i1 = np.empty(4096, dtype=np.int32)
    d1 = np.empty(4096, dtype=np.float64)
    t.KDFindWithinRangeInto(0.8 if k % 2 else 5.0, q[k:k + 1], o1, i1, d1)      # mapped row, both kernels
    t.KDFindNearest(q[k:k + 1])
    t.ExplicitEdgeCheck(s[k:k + 1], e[k:k + 1])
    t.ExplicitNodeCheck(q[k:k + 1])
n = t.tree_size_
lmc = np.random.default_rng(7).uniform(0, 50, n)
t.WriteLMC(0, lmc)
t.ScatterLMC(np.arange(0, n, 5, dtype=np.int32), lmc[::5] * 0.5)
t.ExtendParents(0.7, synth.box_points(300, 6))
t.ExtendRewire(40.0)
a = np.random.default_rng(8).integers(0, n, 60000).astype(np.int32)
b = np.random.default_rng(9).integers(0, n, 60000).astype(np.int32)
t.ExplicitEdgeCheckIds(a[:5000], b[:5000])
offn, idn, _ = t.KDFindWithinRange(0.9, t.positions()[a[:3000]])
t.AppendEdges(np.repeat(a[:3000], np.diff(offn)).astype(np.int32), idn)
t.ObstacleSweep(3)
t.ObstacleSweep(5, other_mask=np.ones(64, dtype=np.uint8))
t.MoveObstacles([1, 2, 3], [[10.0, 10.0], [20.0, 21.0], [30.0, 5.0]], translate_polygon=True)
t.UpdateObstacles([4], used=[0])
t.ExplicitEdgeCheck(s[:3000], e[:3000])
t.SetTimedObstacles(**synth.timed_obstacles(16, 10))
ts, te = synth.timed_segments(5000, 11)
t.ExplicitEdgeCheck2DTimed(ts, te, 0.5)
print("sanitizer pass ok: %d range hits, %d/%d edges collide" % (len(ids), int(v.sum()), len(v)))
