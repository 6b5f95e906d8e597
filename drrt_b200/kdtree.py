"""Host-side mirror of the reference's KDTree interface, batched.

Names and argument meaning follow include/DRRT/kdtree.h:13-181; every method
takes a BATCH of query points (rows) where the reference takes one, and forwards
to the C ABI (include/drrt_gpu.h).  Results come back as numpy arrays; the
`*_dev` variants take and return torch CUDA tensors and leave everything on the
device.  There is no CPU path here.
"""
import ctypes as C
import threading

import numpy as np

from . import capi
from .capi import (EDGE_DUBINS, EDGE_STRAIGHT, METRIC_EUCLID, METRIC_R2S1_WRAP, PI, c_dp, c_i64p,
                   c_ip, c_u8p, check)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _ptr(a, ty):
    return a.ctypes.data_as(ty)


def _dev_ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class _DevView:
    """__cuda_array_interface__ wrapper so torch can adopt a library-owned device buffer"""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr,
                                         "data": (int(ptr), False), "version": 2}


def dev_tensor(ptr, n, typestr, device=None):
    """torch view (no copy) of n items at device pointer ptr; typestr e.g. '<i4', '<i8', '<f8';
    device: index of the GPU that owns the buffer (default: torch's current device)"""
    import torch
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else int(device))
    if n == 0:
        return torch.empty(0, dtype={"<i4": torch.int32, "<i8": torch.int64, "<f8": torch.float64,
                                  "|u1": torch.uint8}[typestr],
                           device=dev)
    return torch.as_tensor(_DevView(ptr, n, typestr), device=dev)


def _torch_stream(device=None):
    """torch's current stream ON THE STORE'S DEVICE (not on torch's current device)"""
    import torch
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class KDTree:
    """KDTree(int d, VectorXi wraps, VectorXd wrapPoints) kdtree.h:33-36 with
    SetDistanceFunction (:47-49) folded into `metric`."""

    def __init__(self, dims=3, wraps=(2,), wrap_points=(2.0 * PI,), metric=METRIC_R2S1_WRAP,
                 device=0, capacity=0):
        self._lib = capi.load()
        self.dimensions_ = dims
        self.metric = metric
        self.device = device
        w = np.asarray(wraps, dtype=np.int32).reshape(-1)
        wp = np.asarray(wrap_points, dtype=np.float64).reshape(-1)
        self.num_wraps_ = len(w)
        h = C.c_void_p()
        check(self._lib.drrt_store_create(device, dims, len(w), w.ctypes.data_as(C.POINTER(C.c_int)),
                                          _ptr(wp, c_dp), metric, int(capacity), C.byref(h)))
        self._h = h
        # batch + fetch is a two-call protocol on the handle: keep the pair together when several
        # host threads share one tree (the reference's callers hold tree_mutex_ for the same reason)
        self._lock = threading.Lock()

    def _check_dev(self, t, dtype, what, min_numel=0, optional=False):
        """A *_dev entry point hands raw data_ptr()s to the C ABI: a tensor of the wrong dtype,
        a strided view or a tensor on another GPU would be read / written as something else."""
        import torch
        if t is None:
            if optional:
                return None
            raise TypeError("%s: tensor required" % what)
        if not torch.is_tensor(t) or not t.is_cuda:
            raise TypeError("%s: CUDA tensor required" % what)
        if t.dtype != dtype:
            raise TypeError("%s: dtype %s required, got %s" % (what, dtype, t.dtype))
        if not t.is_contiguous():
            raise ValueError("%s: contiguous tensor required" % what)
        if t.device.index != self.device:
            raise ValueError("%s: tensor lives on cuda:%s, the store on cuda:%d"
                             % (what, t.device.index, self.device))
        if t.numel() < min_numel:
            raise ValueError("%s: %d elements required, got %d" % (what, min_numel, t.numel()))
        return t

    def _stream(self):
        return _torch_stream(self.device)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.drrt_store_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- KDTree::tree_size_ kdtree.h:23
    @property
    def tree_size_(self):
        n = C.c_int64()
        check(self._lib.drrt_store_size(self._h, C.byref(n)))
        return n.value

    # -- KDTree::KDInsert kdtree.cpp:87-136 (rows inserted in order)
    def KDInsert(self, positions):
        pos = _f64(positions, (-1, 3))
        first = C.c_int32()
        check(self._lib.drrt_store_insert(self._h, pos.shape[0], _ptr(pos, c_dp), C.byref(first)))
        return first.value

    def KDInsert_dev(self, positions):
        """positions: torch.float64 CUDA tensor (n,3) on this store's device."""
        import torch
        self._check_dev(positions, torch.float64, "KDInsert_dev positions")
        assert positions.dim() == 2 and positions.shape[1] == 3
        first = C.c_int32()
        check(self._lib.drrt_store_insert_dev(self._h, positions.shape[0], _dev_ptr(positions),
                                              self._stream(), C.byref(first)))
        return first.value

    def structure(self, first=0, n=None):
        """kd_parent_/kd_child_L_/kd_child_R_/kd_split_ (kdtreenode.h:17-20,29), -1 = none."""
        n = self.tree_size_ - first if n is None else n
        arrs = [np.empty(n, dtype=np.int32) for _ in range(4)]
        check(self._lib.drrt_store_structure(self._h, first, n, *[_ptr(a, c_ip) for a in arrs]))
        return dict(parent=arrs[0], left=arrs[1], right=arrs[2], split=arrs[3])

    def positions(self, first=0, n=None):
        n = self.tree_size_ - first if n is None else n
        pos = np.empty((n, 3), dtype=np.float64)
        check(self._lib.drrt_store_positions(self._h, first, n, _ptr(pos, c_dp)))
        return pos

    # -- KDTree::KDFindWithinRange kdtree.cpp:794-820
    def KDFindWithinRange(self, range_, query_points):
        """-> (offsets int64[nq+1], ids int32[total], dists f64[total]); each row
        sorted by node id.  range_: scalar or per-query array."""
        q = _f64(query_points, (-1, 3))
        nq = q.shape[0]
        rad = None
        r_all = 0.0
        if np.ndim(range_) == 0:
            r_all = float(range_)
        else:
            rad = _f64(range_, (nq,))
        counts = np.zeros(nq, dtype=np.int32)
        total = C.c_int64()
        with self._lock:
            check(self._lib.drrt_range_batch(self._h, nq, _ptr(q, c_dp),
                                             _ptr(rad, c_dp) if rad is not None else None, r_all,
                                             _ptr(counts, c_ip), C.byref(total)))
            ids = np.empty(total.value, dtype=np.int32)
            dist = np.empty(total.value, dtype=np.float64)
            check(self._lib.drrt_range_fetch(self._h, _ptr(ids, c_ip), _ptr(dist, c_dp)))
        off = np.zeros(nq + 1, dtype=np.int64)
        np.cumsum(counts, out=off[1:])
        return off, ids, dist

    def KDFindWithinRangeInto(self, range_, query_points, offsets, ids, dist=None):
        """The same query in one streamed call into caller buffers (numpy arrays or pinned
        CPU torch tensors: offsets int64[nq+1], ids int32[cap], dist f64[cap], f32[cap] -- the
        narrow rows of drrt_range_batch_into_f32 -- or None for the id sets alone).
        -> total hits.  Raises DrrtError(ERR_CAPACITY) when the rows do not fit; offsets
        are complete then, so the retry can be sized."""
        q = _f64(query_points, (-1, 3))
        nq = q.shape[0]
        rad = None
        r_all = 0.0
        if np.ndim(range_) == 0:
            r_all = float(range_)
        else:
            rad = _f64(range_, (nq,))

        def host_ptr(a, ty, np_dtype):
            if a is None:
                return None, 0
            if isinstance(a, np.ndarray):
                assert a.dtype == np_dtype and a.flags.c_contiguous
                return _ptr(a, ty), a.size
            assert not a.is_cuda and a.is_contiguous()
            return C.cast(a.data_ptr(), ty), a.numel()

        op, on = host_ptr(offsets, capi.c_i64p, np.int64)
        ip, cap = host_ptr(ids, c_ip, np.int32)
        # dist: float64 (the reference's JList key, bit-exact) or float32 (narrow rows) or None
        f32 = dist is not None and (dist.dtype == np.float32 if isinstance(dist, np.ndarray)
                                    else str(dist.dtype) == "torch.float32")
        dp, dn = host_ptr(dist, capi.c_fp if f32 else c_dp, np.float32 if f32 else np.float64)
        assert on >= nq + 1 and (dist is None or dn >= cap)
        total = C.c_int64()
        fn = self._lib.drrt_range_batch_into_f32 if f32 else self._lib.drrt_range_batch_into
        with self._lock:
            check(fn(self._h, nq, _ptr(q, c_dp), _ptr(rad, c_dp) if rad is not None else None,
                     r_all, cap, op, ip, dp, C.byref(total)))
        return total.value

    # KDFindMoreWithinRange kdtree.cpp:822-851 is the same search; the caller's
    # list de-duplication (AddToRangeList :670-682) is the caller's to do.
    KDFindMoreWithinRange = KDFindWithinRange

    def KDFindWithinRange_dev(self, range_, query_points):
        """torch in / torch views out (valid until the next range call)."""
        import torch
        q = self._check_dev(query_points, torch.float64, "KDFindWithinRange_dev queries")
        nq = q.numel() // 3
        rad_t = None
        r_all = 0.0
        if torch.is_tensor(range_):
            rad_t = self._check_dev(range_, torch.float64, "KDFindWithinRange_dev radii", nq)
        else:
            r_all = float(range_)
        res = capi.RangeResult()
        check(self._lib.drrt_range_batch_dev(self._h, nq, _dev_ptr(q), _dev_ptr(rad_t),
                                             r_all, self._stream(), C.byref(res)))
        res.device = self.device
        return res

    def range_pack(self):
        """Plain-CSR form of the last KDFindWithinRange_dev result (rows back to back,
        offsets = exclusive prefix sum of the counts); a device-side copy when the
        result was written in slices (res.packed == 0)."""
        res = capi.RangeResult()
        check(self._lib.drrt_range_pack(self._h, self._stream(), C.byref(res)))
        res.device = self.device
        return res

    @staticmethod
    def range_result_tensors(res):
        """(offsets int64[nq+1], ids int32[extent], dist f64[extent]) as torch views of the
        library-owned result of the last KDFindWithinRange_dev; row q is
        [offsets[q], offsets[q] + counts[q]) -- see range_result_counts.  For a plain
        CSR (extent == total) call range_pack() first."""
        d = getattr(res, "device", None)
        return (dev_tensor(res.offsets_dev, res.nq + 1, "<i8", d), dev_tensor(res.ids_dev, res.extent, "<i4", d),
                dev_tensor(res.dist_dev, res.extent, "<f8", d))

    @staticmethod
    def range_result_counts(res):
        return dev_tensor(res.counts_dev, res.nq, "<i4", getattr(res, "device", None))

    # -- KDTree::KDFindNearest kdtree.cpp:264-307
    def KDFindNearest(self, query_points):
        q = _f64(query_points, (-1, 3))
        nq = q.shape[0]
        ids = np.empty(nq, dtype=np.int32)
        dist = np.empty(nq, dtype=np.float64)
        check(self._lib.drrt_nearest_batch(self._h, nq, _ptr(q, c_dp), _ptr(ids, c_ip),
                                           _ptr(dist, c_dp)))
        return ids, dist

    # -- KDTree::KDFindKNearest kdtree.cpp:630-666 (by intent)
    def KDFindKNearest(self, k, query_points):
        q = _f64(query_points, (-1, 3))
        nq = q.shape[0]
        ids = np.empty((nq, k), dtype=np.int32)
        dist = np.empty((nq, k), dtype=np.float64)
        check(self._lib.drrt_knn_batch(self._h, nq, _ptr(q, c_dp), int(k), _ptr(ids, c_ip),
                                       _ptr(dist, c_dp)))
        return ids, dist

    def KDFindKNearest_dev(self, k, q, ids_out, dist_out):
        import torch
        self._check_dev(q, torch.float64, "KDFindKNearest_dev queries")
        nq = q.numel() // 3
        self._check_dev(ids_out, torch.int32, "KDFindKNearest_dev ids_out", nq * int(k))
        self._check_dev(dist_out, torch.float64, "KDFindKNearest_dev dist_out", nq * int(k))
        check(self._lib.drrt_knn_batch_dev(self._h, nq, _dev_ptr(q), int(k),
                                           _dev_ptr(ids_out), _dev_ptr(dist_out), self._stream()))

    # ---- ConfigSpace::obstacles_ + the edge / node checks (obstacle.h:252-282) ----
    def SetObstacles(self, kind, used, life_span, origin, radius, vert_off, verts):
        kind = np.ascontiguousarray(kind, dtype=np.int32)
        used = np.ascontiguousarray(used, dtype=np.uint8)
        life = _f64(life_span)
        origin = _f64(origin, (-1, 2))
        radius = _f64(radius)
        voff = np.ascontiguousarray(vert_off, dtype=np.int32)
        verts = _f64(verts, (-1, 2))
        check(self._lib.drrt_obstacles_set(self._h, len(kind), _ptr(kind, c_ip), _ptr(used, c_u8p),
                                           _ptr(life, c_dp), _ptr(origin, c_dp), _ptr(radius, c_dp),
                                           _ptr(voff, c_ip), _ptr(verts, c_dp)))

    # -- ExplicitEdgeCheck(C, edge) obstacle.cpp:879-923
    def ExplicitEdgeCheck(self, start, end, robot_radius=0.5, r_min=1.0, edge_kind=EDGE_DUBINS,
                          in_warmup=False, want_length=True):
        s = _f64(start, (-1, 3))
        e = _f64(end, (-1, 3))
        n = s.shape[0]
        verdict = np.zeros(n, dtype=np.uint8)
        length = np.zeros(n, dtype=np.float64) if want_length else None
        check(self._lib.drrt_edgecheck_batch(self._h, n, _ptr(s, c_dp), _ptr(e, c_dp), edge_kind,
                                             float(robot_radius), float(r_min), int(in_warmup),
                                             _ptr(verdict, c_u8p),
                                             _ptr(length, c_dp) if want_length else None))
        return (verdict, length) if want_length else verdict

    def ExplicitEdgeCheckIds(self, start_ids, end_ids, robot_radius=0.5, r_min=1.0,
                             edge_kind=EDGE_DUBINS, in_warmup=False):
        si = np.ascontiguousarray(start_ids, dtype=np.int32)
        ei = np.ascontiguousarray(end_ids, dtype=np.int32)
        n = si.shape[0]
        verdict = np.zeros(n, dtype=np.uint8)
        length = np.zeros(n, dtype=np.float64)
        check(self._lib.drrt_edgecheck_ids_batch(self._h, n, _ptr(si, c_ip), _ptr(ei, c_ip),
                                                 edge_kind, float(robot_radius), float(r_min),
                                                 int(in_warmup), _ptr(verdict, c_u8p),
                                                 _ptr(length, c_dp)))
        return verdict, length

    # -- edge->ExplicitEdgeCheck(O) for a subset of the obstacle table: the re-checks of
    #    AddObstacle obstacle.cpp:612-667 / RemoveObstacle :669-754
    def ExplicitEdgeCheckSubset(self, obstacle_mask, start=None, end=None, start_ids=None,
                                end_ids=None, robot_radius=0.5, r_min=1.0, edge_kind=EDGE_DUBINS):
        """obstacle_mask: one byte per obstacle of the table given to SetObstacles (non-zero =
        takes part).  Edges by poses (start/end, n x 3) or by node ids."""
        mask = np.ascontiguousarray(obstacle_mask, dtype=np.uint8)
        s = e = si = ei = None
        if start_ids is not None:
            si = np.ascontiguousarray(start_ids, dtype=np.int32)
            ei = np.ascontiguousarray(end_ids, dtype=np.int32)
            n = si.shape[0]
        else:
            s = _f64(start, (-1, 3))
            e = _f64(end, (-1, 3))
            n = s.shape[0]
        verdict = np.zeros(n, dtype=np.uint8)
        length = np.zeros(n, dtype=np.float64)
        p = lambda a, t: _ptr(a, t) if a is not None else None
        check(self._lib.drrt_edgecheck_subset_batch(self._h, n, p(s, c_dp), p(e, c_dp), p(si, c_ip),
                                                    p(ei, c_ip), edge_kind, float(robot_radius),
                                                    float(r_min), _ptr(mask, c_u8p), len(mask),
                                                    _ptr(verdict, c_u8p), _ptr(length, c_dp)))
        return verdict, length

    # -- FindPointsInConflictWithObstacle obstacle.cpp:540-610 (Dubins space, :554-559)
    def FindPointsInConflictWithObstacle(self, obstacle_indices, robot_radius=0.5):
        """-> (offsets int64[n+1], ids, dists): the nodes within robot_radius + O.radius_ + PI
        of (O.origin_(0), O.origin_(1), PI), one row per listed obstacle."""
        idx = np.ascontiguousarray(obstacle_indices, dtype=np.int32)
        counts = np.zeros(len(idx), dtype=np.int32)
        total = C.c_int64()
        with self._lock:
            check(self._lib.drrt_obstacle_conflict_batch(self._h, len(idx), _ptr(idx, c_ip),
                                                         float(robot_radius), _ptr(counts, c_ip),
                                                         C.byref(total)))
            ids = np.empty(total.value, dtype=np.int32)
            dist = np.empty(total.value, dtype=np.float64)
            check(self._lib.drrt_range_fetch(self._h, _ptr(ids, c_ip), _ptr(dist, c_dp)))
        off = np.zeros(len(idx) + 1, dtype=np.int64)
        np.cumsum(counts, out=off[1:])
        return off, ids, dist

    # -- ExplicitEdgeCheck2D kinds 6 / 7 (obstacle.cpp:805-875): discs moving along (dx, dy, time) paths
    def SetTimedObstacles(self, kind, used, life_span, origin, radius, path_off, path):
        kind = np.ascontiguousarray(kind, dtype=np.int32)
        used = np.ascontiguousarray(used, dtype=np.uint8)
        life = _f64(life_span)
        origin = _f64(origin, (-1, 2))
        radius = _f64(radius)
        poff = np.ascontiguousarray(path_off, dtype=np.int32)
        path = _f64(path, (-1, 3))
        check(self._lib.drrt_timed_obstacles_set(self._h, len(kind), _ptr(kind, c_ip), _ptr(used, c_u8p),
                                                 _ptr(life, c_dp), _ptr(origin, c_dp), _ptr(radius, c_dp),
                                                 _ptr(poff, c_ip), _ptr(path, c_dp)))

    def ExplicitEdgeCheck2DTimed(self, start_xyt, end_xyt, radius=0.5):
        s = _f64(start_xyt, (-1, 3))
        e = _f64(end_xyt, (-1, 3))
        verdict = np.zeros(s.shape[0], dtype=np.uint8)
        check(self._lib.drrt_edgecheck2d_timed_batch(self._h, s.shape[0], _ptr(s, c_dp), _ptr(e, c_dp),
                                                     float(radius), _ptr(verdict, c_u8p)))
        return verdict

    # -- Obstacle::UpdateObstacles / NextOrigin obstacle.cpp:12-22, 182-195: only the moved entries
    def MoveObstacles(self, obstacle_indices, new_origins, translate_polygon=False):
        idx = np.ascontiguousarray(obstacle_indices, dtype=np.int32)
        org = _f64(new_origins, (-1, 2))
        assert len(idx) == len(org)
        check(self._lib.drrt_obstacles_move(self._h, len(idx), _ptr(idx, c_ip), _ptr(org, c_dp),
                                            int(bool(translate_polygon))))

    def UpdateObstacles(self, obstacle_indices, used=None, life_span=None):
        idx = np.ascontiguousarray(obstacle_indices, dtype=np.int32)
        u = np.ascontiguousarray(used, dtype=np.uint8) if used is not None else None
        l = _f64(life_span) if life_span is not None else None
        check(self._lib.drrt_obstacles_update(self._h, len(idx), _ptr(idx, c_ip),
                                              _ptr(u, c_u8p) if u is not None else None,
                                              _ptr(l, c_dp) if l is not None else None))

    # -- the planner's stored neighbour edges on the device + AddObstacle / RemoveObstacle sweeps
    #    (obstacle.cpp:612-754) in one call each
    def AppendEdges(self, start_ids, end_ids):
        a = np.ascontiguousarray(start_ids, dtype=np.int32)
        b = np.ascontiguousarray(end_ids, dtype=np.int32)
        assert a.size == b.size
        first = C.c_int64()
        check(self._lib.drrt_edges_append(self._h, a.size, _ptr(a, c_ip), _ptr(b, c_ip), C.byref(first)))
        return first.value

    @property
    def stored_edges_(self):
        n = C.c_int64()
        check(self._lib.drrt_edges_count(self._h, C.byref(n)))
        return n.value

    def ClearEdges(self):
        check(self._lib.drrt_edges_clear(self._h))

    def ObstacleSweep(self, obstacle_index, other_mask=None, robot_radius=0.5, r_min=1.0,
                      edge_kind=EDGE_DUBINS):
        """-> dict(conflict_nodes, edges_checked, edge_index int64[n], flags uint8[n]): the stored
        edges starting in a node in conflict with the obstacle that collide with it (bit 0), and
        whether they also collide with another obstacle of other_mask (bit 1; RemoveObstacle)."""
        m = np.ascontiguousarray(other_mask, dtype=np.uint8) if other_mask is not None else None
        nc, ne, no = C.c_int64(), C.c_int64(), C.c_int64()
        with self._lock:
            check(self._lib.drrt_obstacle_sweep(self._h, int(obstacle_index),
                                                _ptr(m, c_u8p) if m is not None else None,
                                                len(m) if m is not None else 0, edge_kind,
                                                float(robot_radius), float(r_min), C.byref(nc), C.byref(ne),
                                                C.byref(no)))
            idx = np.empty(no.value, dtype=np.int64)
            fl = np.empty(no.value, dtype=np.uint8)
            check(self._lib.drrt_obstacle_sweep_fetch(self._h, _ptr(idx, c_i64p), _ptr(fl, c_u8p)))
        return dict(conflict_nodes=nc.value, edges_checked=ne.value, edge_index=idx, flags=fl)

    def ExplicitEdgeCheck_dev(self, verdict_out, start=None, end=None, start_ids=None,
                              end_ids=None, robot_radius=0.5, r_min=1.0, edge_kind=EDGE_DUBINS,
                              length_out=None):
        import torch
        self._check_dev(verdict_out, torch.uint8, "ExplicitEdgeCheck_dev verdict_out")
        n = verdict_out.shape[0]
        self._check_dev(start, torch.float64, "ExplicitEdgeCheck_dev start", 3 * n, optional=True)
        self._check_dev(end, torch.float64, "ExplicitEdgeCheck_dev end", 3 * n, optional=True)
        self._check_dev(start_ids, torch.int32, "ExplicitEdgeCheck_dev start_ids", n, optional=True)
        self._check_dev(end_ids, torch.int32, "ExplicitEdgeCheck_dev end_ids", n, optional=True)
        self._check_dev(length_out, torch.float64, "ExplicitEdgeCheck_dev length_out", n, optional=True)
        check(self._lib.drrt_edgecheck_batch_dev(self._h, n, _dev_ptr(start), _dev_ptr(end),
                                                 _dev_ptr(start_ids), _dev_ptr(end_ids), edge_kind,
                                                 float(robot_radius), float(r_min),
                                                 _dev_ptr(verdict_out), _dev_ptr(length_out),
                                                 self._stream()))

    # -- Extend's hot path fused (drrt.cpp:194-362): neighbours, both edges, best parent
    def Extend(self, range_, new_points, lmc=None, robot_radius=0.5, r_min=1.0, in_warmup=False):
        """-> dict(offsets, ids, dist, unsafe[total,2], length[total,2], best_parent, best_cost);
        column 0 = new->near (FindBestParent), column 1 = near->new (Extend's second pass)."""
        q = _f64(new_points, (-1, 3))
        nq = q.shape[0]
        rad, r_all = (None, float(range_)) if np.ndim(range_) == 0 else (_f64(range_, (nq,)), 0.0)
        counts = np.zeros(nq, dtype=np.int32)
        total = C.c_int64()
        best = np.full(nq, -1, dtype=np.int32) if lmc is not None else None
        cost = np.zeros(nq, dtype=np.float64) if lmc is not None else None
        lm = _f64(lmc) if lmc is not None else None
        if lm is not None and lm.size != self.tree_size_:
            # the C side uploads tree_size_ doubles from this array
            raise ValueError("Extend: lmc has %d entries, the store %d nodes" % (lm.size, self.tree_size_))
        with self._lock:
            check(self._lib.drrt_extend_batch(self._h, nq, _ptr(q, c_dp),
                                              _ptr(rad, c_dp) if rad is not None else None, r_all,
                                              float(robot_radius), float(r_min), int(in_warmup),
                                              _ptr(lm, c_dp) if lm is not None else None, _ptr(counts, c_ip),
                                              C.byref(total), _ptr(best, c_ip) if best is not None else None,
                                              _ptr(cost, c_dp) if cost is not None else None))
            t = total.value
            ids = np.empty(t, dtype=np.int32)
            dist = np.empty(t, dtype=np.float64)
            unsafe = np.empty((t, 2), dtype=np.uint8)
            length = np.empty((t, 2), dtype=np.float64)
            check(self._lib.drrt_extend_fetch(self._h, _ptr(ids, c_ip), _ptr(dist, c_dp),
                                              _ptr(unsafe, c_u8p), _ptr(length, c_dp)))
        off = np.zeros(nq + 1, dtype=np.int64)
        np.cumsum(counts, out=off[1:])
        return dict(offsets=off, ids=ids, dist=dist, unsafe=unsafe, length=length, best_parent=best,
                    best_cost=cost)

    # -- device mirror of KDTreeNode::rrt_LMC_ (kdtreenode.h:35-73), read by Extend(lmc=None)
    def WriteLMC(self, first, values):
        v = _f64(values)
        check(self._lib.drrt_lmc_write(self._h, int(first), v.size, _ptr(v, c_dp)))

    def ScatterLMC(self, ids, values):
        i = np.ascontiguousarray(ids, dtype=np.int32)
        v = _f64(values)
        assert i.size == v.size
        check(self._lib.drrt_lmc_scatter(self._h, i.size, _ptr(i, c_ip), _ptr(v, c_dp)))

    def ReadLMC(self, first=0, n=None):
        n = self.tree_size_ - first if n is None else n
        v = np.empty(n, dtype=np.float64)
        check(self._lib.drrt_lmc_read(self._h, int(first), n, _ptr(v, c_dp)))
        return v

    def ExtendParents(self, range_, new_points, robot_radius=0.5, r_min=1.0, in_warmup=False):
        """What FindBestParent needs back on the host (drrt.cpp:367-460) and nothing else: the
        neighbour COUNT, the best parent and its cost per new node, reduced against the device
        mirror of rrt_LMC_ -- the neighbour rows and their 2 x hits edge verdicts never cross PCIe.
        -> (counts int32[nq], best_parent int32[nq], best_cost f64[nq], total)"""
        q = _f64(new_points, (-1, 3))
        nq = q.shape[0]
        rad, r_all = (None, float(range_)) if np.ndim(range_) == 0 else (_f64(range_, (nq,)), 0.0)
        counts = np.zeros(nq, dtype=np.int32)
        best = np.full(nq, -1, dtype=np.int32)
        cost = np.zeros(nq, dtype=np.float64)
        total = C.c_int64()
        with self._lock:
            check(self._lib.drrt_extend_batch(self._h, nq, _ptr(q, c_dp),
                                              _ptr(rad, c_dp) if rad is not None else None, r_all,
                                              float(robot_radius), float(r_min), int(in_warmup), None,
                                              _ptr(counts, c_ip), C.byref(total), _ptr(best, c_ip),
                                              _ptr(cost, c_dp)))
        return counts, best, cost, total.value

    def ExtendRewire(self, move_goal_lmc):
        """Extend's rewiring candidates (drrt.cpp:330-333) of the last Extend / ExtendParents call
        -> (query_index int32[n], near_id int32[n], new_lmc f64[n]) in pair order"""
        n = C.c_int64()
        with self._lock:
            check(self._lib.drrt_extend_rewire(self._h, float(move_goal_lmc), C.byref(n)))
            qi = np.empty(n.value, dtype=np.int32)
            ni = np.empty(n.value, dtype=np.int32)
            lm = np.empty(n.value, dtype=np.float64)
            check(self._lib.drrt_extend_rewire_fetch(self._h, _ptr(qi, c_ip), _ptr(ni, c_ip), _ptr(lm, c_dp)))
        return qi, ni, lm

    def Extend_dev(self, range_, new_points, lmc=None, best_parent=None, best_cost=None,
                   robot_radius=0.5, r_min=1.0):
        """device-resident Extend: torch tensors in; returns (RangeResult rows, unsafe u8[2*total],
        length f64[2*total]) as views of library-owned buffers (valid until the next call)."""
        import torch
        q = self._check_dev(new_points, torch.float64, "Extend_dev new_points")
        nq = q.numel() // 3
        rad_t, r_all = (range_, 0.0) if torch.is_tensor(range_) else (None, float(range_))
        self._check_dev(rad_t, torch.float64, "Extend_dev radii", nq, optional=True)
        self._check_dev(lmc, torch.float64, "Extend_dev lmc", self.tree_size_, optional=True)
        self._check_dev(best_parent, torch.int32, "Extend_dev best_parent", nq, optional=True)
        self._check_dev(best_cost, torch.float64, "Extend_dev best_cost", nq, optional=True)
        res = capi.RangeResult()
        u = C.c_void_p()
        ln = C.c_void_p()
        check(self._lib.drrt_extend_batch_dev(self._h, nq, _dev_ptr(q), _dev_ptr(rad_t), r_all,
                                              float(robot_radius), float(r_min), _dev_ptr(lmc),
                                              self._stream(), C.byref(res), C.byref(u), C.byref(ln),
                                              _dev_ptr(best_parent), _dev_ptr(best_cost)))
        res.device = self.device
        return (res, dev_tensor(u.value, 2 * res.total, "|u1", self.device),
                dev_tensor(ln.value, 2 * res.total, "<f8", self.device))

    # -- LineCheck obstacle.cpp:1092-1143
    def LineCheck(self, node1, node2, robot_radius=0.5):
        return self.ExplicitEdgeCheck(node1, node2, robot_radius=robot_radius,
                                      edge_kind=EDGE_STRAIGHT, want_length=False)

    # -- DubinsEdge::CalculateTrajectory dubinsedge.cpp:172-830
    def CalculateTrajectory(self, start, end, r_min=1.0):
        s = _f64(start, (-1, 3))
        e = _f64(end, (-1, 3))
        n = s.shape[0]
        xy = np.zeros((n, capi.TRAJ_MAX, 2), dtype=np.float64)
        npts = np.zeros(n, dtype=np.int32)
        types = C.create_string_buffer(4 * n)
        length = np.zeros(n, dtype=np.float64)
        check(self._lib.drrt_dubins_trajectory_batch(self._h, n, _ptr(s, c_dp), _ptr(e, c_dp),
                                                     float(r_min), _ptr(xy, c_dp), _ptr(npts, c_ip),
                                                     types, _ptr(length, c_dp)))
        names = [types.raw[4 * i:4 * i + 3].decode() for i in range(n)]
        return xy, npts, names, length

    def DubinsWords(self, start, end, r_min=1.0):
        """edge_type_ and dist_ after CalculateTrajectory, without the polyline (all six candidates
        solved in double: the reference's own sequence)."""
        s = _f64(start, (-1, 3))
        e = _f64(end, (-1, 3))
        n = s.shape[0]
        types = C.create_string_buffer(4 * n)
        length = np.zeros(n, dtype=np.float64)
        check(self._lib.drrt_dubins_trajectory_batch(self._h, n, _ptr(s, c_dp), _ptr(e, c_dp),
                                                     float(r_min), None, None, types, _ptr(length, c_dp)))
        names = np.frombuffer(types.raw, dtype="S4").astype("S3")
        return names, length

    # -- ExplicitNodeCheck obstacle.cpp:1087-1090
    def ExplicitNodeCheck(self, points, robot_radius=0.5, collision_distance=0.1):
        p = _f64(points, (-1, 3))
        verdict = np.zeros(p.shape[0], dtype=np.uint8)
        check(self._lib.drrt_pointcheck_batch(self._h, p.shape[0], _ptr(p, c_dp),
                                              float(robot_radius), float(collision_distance),
                                              _ptr(verdict, c_u8p)))
        return verdict
