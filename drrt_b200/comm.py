"""N GPUs behind one handle in ONE host process (include/drrt_gpu.h, drrt_comm_*).

The reference planner is a single process whose threads share one KDTree; CommKDTree is that
tree over one replica of the node store per GPU: inserts cross PCIe once and reach the other
replicas by ncclBroadcast over NVLink, query / edge batches are sharded into contiguous index
ranges and every GPU copies its rows straight into the caller's buffers.  Method names follow
drrt_b200.KDTree (i.e. include/DRRT/kdtree.h); results are identical for every GPU count.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import EDGE_DUBINS, EDGE_STRAIGHT, METRIC_R2S1_WRAP, PI, c_dp, c_i64p, c_ip, c_u8p, check
from .kdtree import _f64, _ptr


def pinned_empty(n, dtype):
    """numpy view of n items of page-locked host memory from drrt_host_alloc (freed with the
    returned owner object)"""
    dt = np.dtype(dtype)
    p = C.c_void_p()
    check(capi.load().drrt_host_alloc(max(1, int(n)) * dt.itemsize, C.byref(p)))
    buf = (C.c_char * (max(1, int(n)) * dt.itemsize)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dt, count=int(n))

    class _Owner:
        def __init__(self, ptr):
            self.ptr = ptr

        def __del__(self):
            try:
                capi.load().drrt_host_free(self.ptr)
            except Exception:
                pass

    return arr, _Owner(p)


class CommKDTree:
    def __init__(self, n_gpus, devices=None, dims=3, wraps=(2,), wrap_points=(2.0 * PI,),
                 metric=METRIC_R2S1_WRAP, capacity=0):
        self._lib = capi.load()
        w = np.asarray(wraps, dtype=np.int32).reshape(-1)
        wp = np.asarray(wrap_points, dtype=np.float64).reshape(-1)
        dv = None if devices is None else np.asarray(devices, dtype=np.int32)
        h = C.c_void_p()
        check(self._lib.drrt_comm_init(int(n_gpus), dv.ctypes.data_as(C.POINTER(C.c_int)) if dv is not None else None,
                                       dims, len(w), w.ctypes.data_as(C.POINTER(C.c_int)), _ptr(wp, c_dp),
                                       metric, int(capacity), C.byref(h)))
        self._c = h
        self.n_gpus = int(n_gpus)

    def close(self):
        if getattr(self, "_c", None):
            self._lib.drrt_comm_destroy(self._c)
            self._c = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def tree_size_(self):
        n = C.c_int64()
        check(self._lib.drrt_store_size(self._lib.drrt_comm_store(self._c, 0), C.byref(n)))
        return n.value

    @property
    def broadcast_bytes(self):
        return int(self._lib.drrt_comm_broadcast_bytes(self._c))

    def replica_handle(self, rank):
        return C.c_void_p(self._lib.drrt_comm_store(self._c, int(rank)))

    # KDTree::KDInsert on every replica (src/kdtree.cpp:87-136)
    def KDInsert(self, positions):
        pos = _f64(positions, (-1, 3))
        first = C.c_int32()
        check(self._lib.drrt_comm_insert(self._c, pos.shape[0], _ptr(pos, c_dp), C.byref(first)))
        return first.value

    def SetObstacles(self, kind, used, life_span, origin, radius, vert_off, verts):
        kind = np.ascontiguousarray(kind, dtype=np.int32)
        used = np.ascontiguousarray(used, dtype=np.uint8)
        life = _f64(life_span)
        origin = _f64(origin, (-1, 2))
        radius = _f64(radius)
        voff = np.ascontiguousarray(vert_off, dtype=np.int32)
        verts = _f64(verts, (-1, 2))
        check(self._lib.drrt_comm_obstacles_set(self._c, len(kind), _ptr(kind, c_ip), _ptr(used, c_u8p),
                                                _ptr(life, c_dp), _ptr(origin, c_dp), _ptr(radius, c_dp),
                                                _ptr(voff, c_ip), _ptr(verts, c_dp)))

    # KDTree::KDFindWithinRange (src/kdtree.cpp:794-820), query batch sharded over the GPUs
    def KDFindWithinRangeInto(self, range_, query_points, offsets, ids, dist=None):
        """one call into caller buffers (numpy, ideally pinned: pinned_empty); dist float64,
        float32 (narrow rows) or None (id sets alone).  -> total hits"""
        q = _f64(query_points, (-1, 3))
        nq = q.shape[0]
        rad, r_all = (None, float(range_)) if np.ndim(range_) == 0 else (_f64(range_, (nq,)), 0.0)
        assert offsets.dtype == np.int64 and offsets.size >= nq + 1 and ids.dtype == np.int32
        f32 = dist is not None and dist.dtype == np.float32
        assert dist is None or dist.dtype in (np.float32, np.float64)
        assert dist is None or dist.size >= ids.size
        fn = self._lib.drrt_comm_range_batch_into_f32 if f32 else self._lib.drrt_comm_range_batch_into
        dp = None if dist is None else _ptr(dist, capi.c_fp if f32 else c_dp)
        total = C.c_int64()
        check(fn(self._c, nq, _ptr(q, c_dp), _ptr(rad, c_dp) if rad is not None else None, r_all,
                 ids.size, _ptr(offsets, c_i64p), _ptr(ids, c_ip), dp, C.byref(total)))
        return total.value

    def KDFindWithinRange(self, range_, query_points, dist_dtype=np.float64):
        """-> (offsets, ids, dist); sizes the buffers by retrying once on ERR_CAPACITY"""
        q = _f64(query_points, (-1, 3))
        nq = q.shape[0]
        off = np.zeros(nq + 1, dtype=np.int64)
        cap = max(1024, 64 * nq)
        for _ in range(2):
            ids = np.empty(cap, dtype=np.int32)
            dist = np.empty(cap, dtype=dist_dtype) if dist_dtype is not None else None
            try:
                total = self.KDFindWithinRangeInto(range_, q, off, ids, dist)
                return off, ids[:total], (dist[:total] if dist is not None else None)
            except capi.DrrtError as e:
                if e.code != capi.ERR_CAPACITY:
                    raise
                cap = int(off[nq])
        raise RuntimeError("unreachable")

    def KDFindNearest(self, query_points):
        q = _f64(query_points, (-1, 3))
        ids = np.empty(q.shape[0], dtype=np.int32)
        dist = np.empty(q.shape[0], dtype=np.float64)
        check(self._lib.drrt_comm_nearest_batch(self._c, q.shape[0], _ptr(q, c_dp), _ptr(ids, c_ip), _ptr(dist, c_dp)))
        return ids, dist

    def KDFindKNearest(self, k, query_points):
        q = _f64(query_points, (-1, 3))
        ids = np.empty((q.shape[0], k), dtype=np.int32)
        dist = np.empty((q.shape[0], k), dtype=np.float64)
        check(self._lib.drrt_comm_knn_batch(self._c, q.shape[0], _ptr(q, c_dp), int(k), _ptr(ids, c_ip),
                                            _ptr(dist, c_dp)))
        return ids, dist

    # ExplicitEdgeCheck(C, edge) src/obstacle.cpp:879-923 over a sharded edge batch
    def ExplicitEdgeCheck(self, start, end, robot_radius=0.5, r_min=1.0, edge_kind=EDGE_DUBINS,
                          in_warmup=False, verdict=None, length=None):
        s = _f64(start, (-1, 3))
        e = _f64(end, (-1, 3))
        n = s.shape[0]
        verdict = np.zeros(n, dtype=np.uint8) if verdict is None else verdict
        length = np.zeros(n, dtype=np.float64) if length is None else length
        check(self._lib.drrt_comm_edgecheck_batch(self._c, n, _ptr(s, c_dp), _ptr(e, c_dp), edge_kind,
                                                  float(robot_radius), float(r_min), int(in_warmup),
                                                  _ptr(verdict, c_u8p), _ptr(length, c_dp)))
        return verdict, length

    def ExplicitEdgeCheckIds(self, start_ids, end_ids, robot_radius=0.5, r_min=1.0, edge_kind=EDGE_DUBINS,
                             in_warmup=False):
        si = np.ascontiguousarray(start_ids, dtype=np.int32)
        ei = np.ascontiguousarray(end_ids, dtype=np.int32)
        verdict = np.zeros(si.size, dtype=np.uint8)
        length = np.zeros(si.size, dtype=np.float64)
        check(self._lib.drrt_comm_edgecheck_ids_batch(self._c, si.size, _ptr(si, c_ip), _ptr(ei, c_ip), edge_kind,
                                                      float(robot_radius), float(r_min), int(in_warmup),
                                                      _ptr(verdict, c_u8p), _ptr(length, c_dp)))
        return verdict, length

    def LineCheck(self, node1, node2, robot_radius=0.5):
        return self.ExplicitEdgeCheck(node1, node2, robot_radius=robot_radius, edge_kind=EDGE_STRAIGHT)[0]

    def ExplicitNodeCheck(self, points, robot_radius=0.5, collision_distance=0.1):
        p = _f64(points, (-1, 3))
        verdict = np.zeros(p.shape[0], dtype=np.uint8)
        check(self._lib.drrt_comm_pointcheck_batch(self._c, p.shape[0], _ptr(p, c_dp), float(robot_radius),
                                                   float(collision_distance), _ptr(verdict, c_u8p)))
        return verdict

    def WriteLMC(self, first, values):
        v = _f64(values)
        check(self._lib.drrt_comm_lmc_write(self._c, int(first), v.size, _ptr(v, c_dp)))

    def ScatterLMC(self, ids, values):
        i = np.ascontiguousarray(ids, dtype=np.int32)
        v = _f64(values)
        check(self._lib.drrt_comm_lmc_scatter(self._c, i.size, _ptr(i, c_ip), _ptr(v, c_dp)))

    def ExtendParents(self, range_, new_points, lmc=None, robot_radius=0.5, r_min=1.0, in_warmup=False):
        q = _f64(new_points, (-1, 3))
        nq = q.shape[0]
        rad, r_all = (None, float(range_)) if np.ndim(range_) == 0 else (_f64(range_, (nq,)), 0.0)
        counts = np.zeros(nq, dtype=np.int32)
        best = np.full(nq, -1, dtype=np.int32)
        cost = np.zeros(nq, dtype=np.float64)
        lm = _f64(lmc) if lmc is not None else None
        if lm is not None and lm.size != self.tree_size_:
            raise ValueError("ExtendParents: lmc has %d entries, the store %d nodes" % (lm.size, self.tree_size_))
        total = C.c_int64()
        check(self._lib.drrt_comm_extend_batch(self._c, nq, _ptr(q, c_dp),
                                               _ptr(rad, c_dp) if rad is not None else None, r_all,
                                               float(robot_radius), float(r_min), int(in_warmup),
                                               _ptr(lm, c_dp) if lm is not None else None, _ptr(counts, c_ip),
                                               C.byref(total), _ptr(best, c_ip), _ptr(cost, c_dp)))
        return counts, best, cost, total.value
