"""Multi-GPU plumbing for the hot path, ONE PROCESS PER GPU (the one-process
form is drrt_b200.comm / drrt_comm_* in the C ABI).

The node store and obstacle table are REPLICATED -- rank `src` (where the host planner lives)
broadcasts each insert batch / obstacle table over torch.distributed (NCCL over NVLink on
GPUs) and every rank applies it to its local store straight from the received device tensor.
Query and edge batches are SHARDED into contiguous index ranges, one per rank, with no
collective on the data path; the rows are then GATHERED TO `src` ONLY, in global index order
(point-to-point sends of exactly each shard's size: nothing is padded and nothing is delivered
to ranks that do not consume it), so answers do not depend on the number of ranks.

The local back end is any object with the KDTree host interface (drrt_b200.KDTree on a GPU
box; when it also offers the `*_dev` methods and the backend is NCCL, rows stay on the device
from the kernel to the gather).  Collectives use CUDA tensors under NCCL and CPU tensors under
gloo.
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


def _as_tensor(a, dev):
    if torch.is_tensor(a):
        return a if a.device == dev else a.to(dev)
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def broadcast_tensor(arr, src=0, dtype=np.float64):
    """array held by rank src -> the same tensor on every rank's device (shape travels first)"""
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
    return t


def broadcast_array(arr, src=0, dtype=np.float64):
    return broadcast_tensor(arr, src, dtype).cpu().numpy()


def gather_rows(t, dst=0):
    """Concatenation over ranks (rank order) of tensors with different leading sizes, delivered
    to rank dst ONLY (None elsewhere).  The sizes travel in one tiny all_gather; the payload in
    one send per rank of exactly its size."""
    dev = _device()
    world, rank = dist.get_world_size(), dist.get_rank()
    t = _as_tensor(t, dev).contiguous()
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    if rank != dst:
        if sizes[rank]:
            dist.send(t, dst)
        return None
    out = torch.empty((sum(sizes),) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
    at = 0
    for r, s in enumerate(sizes):
        if s:
            if r == rank:
                out[at:at + s] = t
            else:
                dist.recv(out[at:at + s], r)
        at += s
    return out


def _host(t):
    return None if t is None else t.cpu().numpy()


class ReplicatedKDTree:
    """The reference's KDTree interface over N replicas of the node store.  Every rank passes
    the same global batch to the query methods; rank `src` gets the global answer, the other
    ranks get None."""

    def __init__(self, local_tree, src=0):
        self.local = local_tree
        self.src = src
        self.rank = dist.get_rank()
        self.world = dist.get_world_size()
        self.on_device = dist.get_backend() == "nccl" and hasattr(local_tree, "KDFindWithinRange_dev")

    # KDTree::KDInsert on every replica: NCCL broadcast of the insert batch, consumed in place
    def KDInsert(self, positions=None):
        t = broadcast_tensor(positions if self.rank == self.src else None, self.src).reshape(-1, 3)
        if self.on_device:
            return self.local.KDInsert_dev(t.contiguous())
        return self.local.KDInsert(t.cpu().numpy())

    def SetObstacles(self, table=None):
        keys = ("kind", "used", "life_span", "origin", "radius", "vert_off", "verts")
        dtypes = (np.int32, np.uint8, np.float64, np.float64, np.float64, np.int32, np.float64)
        got = {}
        for k, dt in zip(keys, dtypes):
            got[k] = broadcast_array(table[k] if self.rank == self.src else None, self.src, dt)
        self.local.SetObstacles(**got)

    # KDTree::KDFindWithinRange, query batch sharded by contiguous index range
    def KDFindWithinRange(self, range_, query_points):
        q = np.ascontiguousarray(query_points, dtype=np.float64).reshape(-1, 3)
        lo, hi = shard_bounds(len(q), self.rank, self.world)
        r = range_ if np.ndim(range_) == 0 else np.asarray(range_, dtype=np.float64)[lo:hi]
        if self.on_device:
            dev = _device()
            rq = torch.from_numpy(np.ascontiguousarray(q[lo:hi])).to(dev)
            rr = r if np.ndim(r) == 0 else torch.from_numpy(np.ascontiguousarray(r)).to(dev)
            self.local.KDFindWithinRange_dev(rr, rq)
            res = self.local.range_pack()
            _, ids_t, dst_t = self.local.range_result_tensors(res)
            counts_t = self.local.range_result_counts(res).to(torch.int64)
            ids_t, dst_t = ids_t[:res.total], dst_t[:res.total]
        else:
            off, ids_t, dst_t = self.local.KDFindWithinRange(r, q[lo:hi])
            counts_t = np.diff(off).astype(np.int64)
        counts = gather_rows(counts_t, self.src)
        ids = gather_rows(ids_t, self.src)
        dst = gather_rows(dst_t, self.src)
        if self.rank != self.src:
            return None
        goff = torch.zeros(len(q) + 1, dtype=torch.int64, device=counts.device)
        torch.cumsum(counts, 0, out=goff[1:])
        return _host(goff), _host(ids), _host(dst)

    def _sharded(self, fn, query_points):
        q = np.ascontiguousarray(query_points, dtype=np.float64).reshape(-1, 3)
        lo, hi = shard_bounds(len(q), self.rank, self.world)
        outs = fn(q[lo:hi])
        got = [gather_rows(o, self.src) for o in outs]
        return None if self.rank != self.src else tuple(_host(g) for g in got)

    def KDFindNearest(self, query_points):
        return self._sharded(self.local.KDFindNearest, query_points)

    def KDFindKNearest(self, k, query_points):
        return self._sharded(lambda q: self.local.KDFindKNearest(k, q), query_points)

    # ExplicitEdgeCheck over a sharded edge batch
    def ExplicitEdgeCheck(self, start, end, **kw):
        s = np.ascontiguousarray(start, dtype=np.float64).reshape(-1, 3)
        e = np.ascontiguousarray(end, dtype=np.float64).reshape(-1, 3)
        lo, hi = shard_bounds(len(s), self.rank, self.world)
        v, ln = self.local.ExplicitEdgeCheck(s[lo:hi], e[lo:hi], **kw)
        gv, gl = gather_rows(v, self.src), gather_rows(ln, self.src)
        return None if self.rank != self.src else (_host(gv), _host(gl))
