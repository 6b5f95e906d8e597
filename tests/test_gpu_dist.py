"""The same replicated/sharded logic on real GPUs: torchrun, NCCL, world size =
number of visible GPUs (skipped below 2)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
import drrt_b200
from drrt_b200 import dist as ddist, synth
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank = dist.get_rank()
rep = ddist.ReplicatedKDTree(drrt_b200.KDTree(device=local), src=0)
pos = synth.box_points(60000, 301)
rep.KDInsert(pos if rank == 0 else None)
q = synth.box_points(1001, 302)
assert rep.on_device
rows = rep.KDFindWithinRange(1.5, q)
big = rep.KDFindWithinRange(6.0, q[:97])        # CTA-per-query kernel, sliced rows packed before the gather
near = rep.KDFindNearest(q)
rep.SetObstacles(synth.obstacles(64, 303) if rank == 0 else None)
s, e = synth.random_edges(5001, 304)
edges = rep.ExplicitEdgeCheck(s, e)
if rank == 0:
    one = drrt_b200.KDTree(device=0); one.KDInsert(pos); one.SetObstacles(**synth.obstacles(64, 303))
    for got, want in ((rows, one.KDFindWithinRange(1.5, q)), (big, one.KDFindWithinRange(6.0, q[:97])),
                      (near, one.KDFindNearest(q)), (edges, one.ExplicitEdgeCheck(s, e))):
        assert len(got) == len(want)
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
    print("DIST_OK", dist.get_world_size())
else:
    assert rows is None and big is None and near is None and edges is None
dist.destroy_process_group()
'''


def test_nccl_replicated_sharded(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                          "--nproc-per-node", str(min(n, 8)), "--master-addr", "127.0.0.1",
                          "--master-port", "29533", str(script)], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "DIST_OK" in out.stdout


# ---- ONE process, N GPUs through the C ABI (drrt_comm_*) -----------------------------------

def _ngpu():
    import torch
    return torch.cuda.device_count()


def test_comm_one_gpu_equals_store(drrt, po):
    """drrt_comm with a single GPU is the single-store path (runs on any GPU box)"""
    from drrt_b200 import synth
    from drrt_b200.comm import CommKDTree
    c = CommKDTree(1)
    pos = synth.box_points(30000, 311)
    assert c.KDInsert(pos[:20000]) == 0 and c.KDInsert(pos[20000:]) == 20000
    ref = po.RefTree()
    ref.insert(pos)
    q = synth.box_points(700, 312)
    off, ids, dst = c.KDFindWithinRange(2.0, q)
    roff, rids, rdst = ref.range_batch(q, 2.0)
    assert np.array_equal(off, roff) and np.array_equal(ids, rids) and np.array_equal(dst, rdst)
    assert c.broadcast_bytes == 0


def _comm_vs_single(drrt, n_gpus):
    from drrt_b200 import synth
    from drrt_b200.comm import CommKDTree, pinned_empty
    c = CommKDTree(n_gpus, capacity=120_000)
    one = drrt.KDTree(device=0)
    pos = synth.box_points(120_000, 321)
    for lo in range(0, len(pos), 50_000):          # three insert batches, the last one short
        assert c.KDInsert(pos[lo:lo + 50_000]) == lo
        one.KDInsert(pos[lo:lo + 50_000])
    assert c.tree_size_ == len(pos)
    assert c.broadcast_bytes == 24 * len(pos) * (n_gpus - 1)      # NCCL carried every replica's copy
    # every replica holds the reference's pointer tree
    want = one.structure()
    for r in range(n_gpus):
        t = drrt.KDTree.__new__(drrt.KDTree)
        t._lib, t._h, t.device = c._lib, c.replica_handle(r), r
        got = drrt.KDTree.structure(t)
        for k in want:
            assert np.array_equal(got[k], want[k]), (r, k)
        t._h = None                                  # owned by the comm
    q = synth.box_points(1001, 322)                  # odd count: uneven shards
    rad = np.linspace(0.4, 3.5, len(q))              # warp- and CTA-kernel rows in one batch
    want = one.KDFindWithinRange(rad, q)
    got = c.KDFindWithinRange(rad, q)
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    # narrow rows and id sets alone: the same offsets and ids, float-rounded distances
    off32, ids32, d32 = c.KDFindWithinRange(rad, q, dist_dtype=np.float32)
    assert np.array_equal(off32, want[0]) and np.array_equal(ids32, want[1])
    assert np.array_equal(d32, want[2].astype(np.float32))
    offi, idsi, none = c.KDFindWithinRange(rad, q, dist_dtype=None)
    assert none is None and np.array_equal(offi, want[0]) and np.array_equal(idsi, want[1])
    # caller-owned pinned buffers, exact capacity; one too small -> ERR_CAPACITY with the sizes
    total = int(want[0][-1])
    off_p, k1 = pinned_empty(len(q) + 1, np.int64)
    ids_p, k2 = pinned_empty(total, np.int32)
    dst_p, k3 = pinned_empty(total, np.float64)
    assert c.KDFindWithinRangeInto(rad, q, off_p, ids_p, dst_p) == total
    assert np.array_equal(off_p, want[0]) and np.array_equal(ids_p, want[1]) and np.array_equal(dst_p, want[2])
    with pytest.raises(drrt.DrrtError) as ei:
        c.KDFindWithinRangeInto(rad, q, off_p, ids_p[:total - 1], dst_p[:total - 1])
    assert ei.value.code == drrt.capi.ERR_CAPACITY and off_p[-1] == total
    # fewer queries than GPUs, and none
    for m in (1, 0):
        a = c.KDFindWithinRange(2.0, q[:m])
        b = one.KDFindWithinRange(2.0, q[:m])
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    for a, b in zip(c.KDFindNearest(q), one.KDFindNearest(q)):
        assert np.array_equal(a, b)
    for a, b in zip(c.KDFindKNearest(5, q), one.KDFindKNearest(5, q)):
        assert np.array_equal(a, b)
    o = synth.obstacles(64, 323)
    c.SetObstacles(**o)
    one.SetObstacles(**o)
    s, e = synth.random_edges(2_500_001, 324)        # above the pipelining threshold on one GPU
    for kind in (drrt.EDGE_DUBINS, drrt.EDGE_STRAIGHT):
        v, ln = c.ExplicitEdgeCheck(s, e, edge_kind=kind)
        v1, l1 = one.ExplicitEdgeCheck(s, e, edge_kind=kind)
        assert np.array_equal(v, v1) and np.array_equal(ln, l1)
    a = np.random.default_rng(5).integers(0, len(pos), 40_001).astype(np.int32)
    b = np.random.default_rng(6).integers(0, len(pos), 40_001).astype(np.int32)
    for x, y in zip(c.ExplicitEdgeCheckIds(a, b), one.ExplicitEdgeCheckIds(a, b)):
        assert np.array_equal(x, y)
    pts = synth.box_points(30_001, 325)
    assert np.array_equal(c.ExplicitNodeCheck(pts), one.ExplicitNodeCheck(pts))
    # fused Extend against the replicated rrt_LMC_ mirror: whole array, then a delta update
    lmc = np.random.default_rng(7).uniform(0, 80, len(pos))
    qn = q[:333]
    want = one.Extend(0.9, qn, lmc=lmc)
    cnt, best, cost, total = c.ExtendParents(0.9, qn, lmc=lmc)
    assert np.array_equal(cnt, np.diff(want["offsets"])) and total == len(want["ids"])
    assert np.array_equal(best, want["best_parent"]) and np.array_equal(cost, want["best_cost"])
    ids_ch = np.arange(0, len(pos), 7, dtype=np.int32)
    lmc[ids_ch] *= 0.25
    c.ScatterLMC(ids_ch, lmc[ids_ch])
    want = one.Extend(0.9, qn, lmc=lmc)
    cnt, best, cost, total = c.ExtendParents(0.9, qn)          # lmc=None: the mirror
    assert np.array_equal(best, want["best_parent"]) and np.array_equal(cost, want["best_cost"])
    c.close()


def test_comm_two_gpus_one_process(drrt):
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    _comm_vs_single(drrt, 2)


def test_comm_all_gpus_one_process(drrt):
    n = _ngpu()
    if n < 3:
        pytest.skip("needs >= 3 GPUs (the 2-GPU case is its own test)")
    _comm_vs_single(drrt, min(n, 8))


def test_two_stores_two_devices_two_threads(drrt, po):
    """Two independent handles on two devices in one process, entered from two host threads
    at once: per-device kernel attributes (the > 48 KB shared-memory opt-in of range_cta_kernel)
    and the persistent grid size belong to each store's own device."""
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    import threading
    from drrt_b200 import synth
    pos = synth.box_points(40_000, 331)
    q = synth.box_points(300, 332)
    ref = po.RefTree()
    ref.insert(pos)
    want = ref.range_batch(q, 6.0)                   # ~2300 hits per query: the CTA kernel
    out = {}

    def run(dev):
        t = drrt.KDTree(device=dev)                  # device 1 first: it must not inherit device 0's state
        t.KDInsert(pos)
        for _ in range(3):
            out[dev] = t.KDFindWithinRange(6.0, q)

    th = [threading.Thread(target=run, args=(d,)) for d in (1, 0)]
    [t.start() for t in th]
    [t.join() for t in th]
    for dev in (0, 1):
        for a, b in zip(out[dev], want):
            assert np.array_equal(a, b), dev
