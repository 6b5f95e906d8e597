"""World-size-2 tests of the multi-GPU host logic (drrt_b200/dist.py) on CPU
with the gloo backend: replication by broadcast, contiguous sharding, gather to the
planner's rank in global index order.  The per-rank back end here is the ORACLE wrapped in the
KDTree method names (test infrastructure); on a GPU box it is drrt_b200.KDTree
(tests/test_gpu_dist.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class OracleBackend:
    """oracle behind the drrt_b200.KDTree host interface"""

    def __init__(self):
        from oracle import pyoracle as po
        self.po = po
        self.tree = po.RefTree()
        self.obs = None

    def KDInsert(self, pos):
        return self.tree.insert(pos)

    def KDFindWithinRange(self, r, q):
        return self.tree.range_batch(q, r, nthreads=1)

    def KDFindNearest(self, q):
        return self.tree.nearest_batch(q, nthreads=1)

    def KDFindKNearest(self, k, q):
        return self.tree.knn_batch(q, k, nthreads=1)

    def SetObstacles(self, **o):
        self.obs = self.po.RefObstacles(**o)

    def ExplicitEdgeCheck(self, s, e, **kw):
        return self.obs.edge_check_batch(s, e, nthreads=1)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from drrt_b200 import dist as ddist, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rep = ddist.ReplicatedKDTree(OracleBackend(), src=0)
        pos = synth.box_points(6000, 201)
        # only the source rank holds the insert batches; two batches, second uneven
        assert rep.KDInsert(pos[:4001] if rank == 0 else None) == 0
        assert rep.KDInsert(pos[4001:] if rank == 0 else None) == 4001
        assert rep.local.tree.size == 6000
        q = synth.box_points(101, 202)           # odd count: shards of 51 and 50
        rad = np.linspace(0.5, 2.5, 101)
        rows = rep.KDFindWithinRange(rad, q)
        near = rep.KDFindNearest(q)
        knn = rep.KDFindKNearest(3, q)
        rep.SetObstacles(synth.obstacles(16, 203) if rank == 0 else None)
        s, e = synth.random_edges(333, 204, max_len=2.0)
        edges = rep.ExplicitEdgeCheck(s, e)
        if rank == 0:      # the planner's rank gets the global answer ...
            np.savez(os.path.join(out_dir, "rank%d.npz" % rank), off=rows[0], ids=rows[1], dst=rows[2],
                     nid=near[0], kid=knn[0], v=edges[0], ln=edges[1])
        else:              # ... and nothing is delivered to ranks that do not consume it
            assert rows is None and near is None and knn is None and edges is None
            open(os.path.join(out_dir, "rank%d.none" % rank), "w").close()
    finally:
        dist.destroy_process_group()


def test_shard_bounds():
    from drrt_b200.dist import shard_bounds
    for n in (0, 1, 7, 64, 65536, 10_000_001):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_replicated_sharded_world2(tmp_path, po):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from drrt_b200 import synth
    single = OracleBackend()
    single.KDInsert(synth.box_points(6000, 201))
    q = synth.box_points(101, 202)
    rad = np.linspace(0.5, 2.5, 101)
    off, ids, dst = single.KDFindWithinRange(rad, q)
    nid, _ = single.KDFindNearest(q)
    kid, _ = single.KDFindKNearest(3, q)
    single.SetObstacles(**synth.obstacles(16, 203))
    s, e = synth.random_edges(333, 204, max_len=2.0)
    v, ln = single.ExplicitEdgeCheck(s, e)
    assert os.path.exists(os.path.join(str(tmp_path), "rank1.none"))
    got = np.load(os.path.join(str(tmp_path), "rank0.npz"))
    # identical to the single-process answer
    assert np.array_equal(got["off"], off) and np.array_equal(got["ids"], ids)
    assert np.array_equal(got["dst"], dst)
    assert np.array_equal(got["nid"], nid) and np.array_equal(got["kid"], kid)
    assert np.array_equal(got["v"], v) and np.array_equal(got["ln"], ln)
