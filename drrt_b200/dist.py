"""Multi-GPU plumbing for the hot path: one process per GPU.

The node store and obstacle table are REPLICATED -- rank `src` (where the host
planner lives) broadcasts each insert batch / obstacle table over
torch.distributed (NCCL over NVLink on GPUs), every rank applies it to its local
store.  Query and edge batches are SHARDED into contiguous index ranges, one
per rank, with no collective on the data path; results are gathered back in
global index order, so answers do not depend on the number of ranks.

The local back end is any object with the KDTree host interface
(drrt_b200.KDTree on a GPU box).  Collectives use CUDA tensors under NCCL and
CPU tensors under gloo.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """contiguous [lo, hi) of item range n owned by `rank` (sizes differ by <= 1)"""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _device():
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def broadcast_array(arr, src=0, dtype=np.float64):
    """numpy array from rank src to every rank (shape travels first)"""
    dev = _device()
    rank = dist.get_rank()
    if rank == src:
        arr = np.ascontiguousarray(arr, dtype=dtype)
        meta = torch.tensor([arr.ndim] + list(arr.shape) + [0] * (7 - arr.ndim), dtype=torch.int64)
    else:
        meta = torch.zeros(8, dtype=torch.int64)
    meta = meta.to(dev)
    dist.broadcast(meta, src=src)
    meta = meta.cpu().tolist()
    shape = tuple(meta[1:1 + meta[0]])
    if rank == src:
        t = torch.from_numpy(arr).to(dev)
    else:
        t = torch.empty(shape, dtype=torch.from_numpy(np.empty(0, dtype=dtype)).dtype, device=dev)
    if t.numel():
        dist.broadcast(t, src=src)
    return t.cpu().numpy()


def all_gather_varlen(arr):
    """concatenation over ranks (rank order) of 1-D/2-D arrays with different leading sizes"""
    dev = _device()
    world = dist.get_world_size()
    arr = np.ascontiguousarray(arr)
    n = torch.tensor([arr.shape[0]], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes) if sizes else 0
    pad = np.zeros((mx,) + arr.shape[1:], dtype=arr.dtype)
    pad[:arr.shape[0]] = arr
    t = torch.from_numpy(pad).to(dev)
    outs = [torch.empty_like(t) for _ in range(world)]
    if mx:
        dist.all_gather(outs, t)
    return np.concatenate([o.cpu().numpy()[:s] for o, s in zip(outs, sizes)], axis=0)


class ReplicatedKDTree:
    """The reference's KDTree interface over N replicas of the node store."""

    def __init__(self, local_tree, src=0):
        self.local = local_tree
        self.src = src
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()

    # KDTree::KDInsert on every replica: NCCL broadcast of the insert batch
    def KDInsert(self, positions=None):
        pos = broadcast_array(positions if self.rank == self.src else None, self.src)
        return self.local.KDInsert(pos)

    def SetObstacles(self, table=None):
        keys = ("kind", "used", "life_span", "origin", "radius", "vert_off", "verts")
        dtypes = (np.int32, np.uint8, np.float64, np.float64, np.float64, np.int32, np.float64)
        got = {}
        for k, dt in zip(keys, dtypes):
            got[k] = broadcast_array(table[k] if self.rank == self.src else None, self.src, dt)
        self.local.SetObstacles(**got)

    # KDTree::KDFindWithinRange, query batch sharded by contiguous index range
    def KDFindWithinRange(self, range_, query_points):
        """every rank passes the same global batch; returns the global CSR"""
        q = np.ascontiguousarray(query_points, dtype=np.float64).reshape(-1, 3)
        lo, hi = shard_bounds(len(q), self.rank, self.world)
        r = range_ if np.ndim(range_) == 0 else np.asarray(range_, dtype=np.float64)[lo:hi]
        off, ids, dst = self.local.KDFindWithinRange(r, q[lo:hi])
        counts = all_gather_varlen(np.diff(off).astype(np.int64))
        ids = all_gather_varlen(ids)
        dst = all_gather_varlen(dst)
        goff = np.zeros(len(q) + 1, dtype=np.int64)
        np.cumsum(counts, out=goff[1:])
        return goff, ids, dst

    def KDFindNearest(self, query_points):
        q = np.ascontiguousarray(query_points, dtype=np.float64).reshape(-1, 3)
        lo, hi = shard_bounds(len(q), self.rank, self.world)
        ids, dst = self.local.KDFindNearest(q[lo:hi])
        return all_gather_varlen(ids), all_gather_varlen(dst)

    def KDFindKNearest(self, k, query_points):
        q = np.ascontiguousarray(query_points, dtype=np.float64).reshape(-1, 3)
        lo, hi = shard_bounds(len(q), self.rank, self.world)
        ids, dst = self.local.KDFindKNearest(k, q[lo:hi])
        return all_gather_varlen(ids), all_gather_varlen(dst)

    # ExplicitEdgeCheck over a sharded edge batch
    def ExplicitEdgeCheck(self, start, end, **kw):
        s = np.ascontiguousarray(start, dtype=np.float64).reshape(-1, 3)
        e = np.ascontiguousarray(end, dtype=np.float64).reshape(-1, 3)
        lo, hi = shard_bounds(len(s), self.rank, self.world)
        v, ln = self.local.ExplicitEdgeCheck(s[lo:hi], e[lo:hi], **kw)
        return all_gather_varlen(v), all_gather_varlen(ln)
