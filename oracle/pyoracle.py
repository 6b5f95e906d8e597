"""ctypes front end of the CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  The product package
(drrt_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

METRIC_R2S1 = 0
METRIC_EUCLID = 1
REF_PI = 3.1415926536
REF_INF = 1000000000000.0
TRAJ_MAX = 200

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_u8p = C.POINTER(C.c_ubyte)
c_llp = C.POINTER(C.c_longlong)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"],
                              stdout=subprocess.DEVNULL)
    return so


class _Obstacles(C.Structure):
    _fields_ = [("n", C.c_int), ("kind", c_ip), ("used", c_u8p),
                ("life_span", c_dp), ("origin", c_dp), ("radius", c_dp),
                ("vert_off", c_ip), ("verts", c_dp)]


class _TimedObstacles(C.Structure):
    _fields_ = [("n", C.c_int), ("kind", c_ip), ("used", c_u8p), ("life_span", c_dp), ("origin", c_dp),
                ("radius", c_dp), ("path_off", c_ip), ("path", c_dp)]


class _Traj(C.Structure):
    _fields_ = [("type", C.c_char * 4), ("dist", C.c_double), ("npts", C.c_int),
                ("x", C.c_double * TRAJ_MAX), ("y", C.c_double * TRAJ_MAX)]


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.ref_tree_create.restype = C.c_void_p
        L.ref_tree_create.argtypes = [C.c_int, C.c_int, c_ip, c_dp, C.c_int]
        L.ref_tree_destroy.argtypes = [C.c_void_p]
        L.ref_tree_size.argtypes = [C.c_void_p]
        L.ref_tree_insert.argtypes = [C.c_void_p, C.c_int, c_dp]
        L.ref_tree_structure.argtypes = [C.c_void_p, c_ip, c_ip, c_ip, c_ip]
        L.ref_range.restype = C.c_long
        L.ref_range.argtypes = [C.c_void_p, c_dp, C.c_double, c_ip, c_dp, C.c_long]
        L.ref_range_brute.restype = C.c_long
        L.ref_range_brute.argtypes = [C.c_void_p, c_dp, C.c_double, c_ip, c_dp, C.c_long]
        L.ref_range_batch_count.argtypes = [C.c_void_p, C.c_long, c_dp, c_dp, c_ip, C.c_int]
        L.ref_range_batch_fill.argtypes = [C.c_void_p, C.c_long, c_dp, c_dp, c_llp,
                                           c_ip, c_dp, C.c_int, C.c_int]
        L.ref_range_last_visits.restype = C.c_long
        L.ref_nearest_batch.argtypes = [C.c_void_p, C.c_long, c_dp, c_ip, c_dp, C.c_int]
        L.ref_knn_batch.argtypes = [C.c_void_p, C.c_long, c_dp, C.c_int, c_ip, c_dp, C.c_int]
        L.ref_ghosts.argtypes = [C.c_void_p, c_dp, C.c_double, c_dp, C.c_int]
        L.ref_metric.restype = C.c_double
        L.ref_metric.argtypes = [C.c_int, c_dp, c_dp]
        L.ref_dist_sqrd_point_segment.restype = C.c_double
        L.ref_dist_sqrd_point_segment.argtypes = [c_dp, c_dp, c_dp]
        L.ref_segment_dist_sqrd.restype = C.c_double
        L.ref_segment_dist_sqrd.argtypes = [c_dp, c_dp, c_dp, c_dp]
        L.ref_right_turn_dist.restype = C.c_double
        L.ref_right_turn_dist.argtypes = [c_dp, c_dp, c_dp, C.c_double]
        L.ref_left_turn_dist.restype = C.c_double
        L.ref_left_turn_dist.argtypes = [c_dp, c_dp, c_dp, C.c_double]
        L.ref_point_in_polygon.argtypes = [c_dp, c_dp, C.c_int]
        L.ref_dist_to_polygon_sqrd.restype = C.c_double
        L.ref_dist_to_polygon_sqrd.argtypes = [c_dp, c_dp, C.c_int]
        L.ref_edge_check_2d.argtypes = [C.POINTER(_Obstacles), C.c_int, c_dp, c_dp,
                                        C.c_double, c_dp]
        L.ref_dubins_trajectory.argtypes = [c_dp, c_dp, C.c_double, C.c_int, C.POINTER(_Traj)]
        L.ref_line_trajectory.argtypes = [c_dp, c_dp, C.POINTER(_Traj)]
        L.ref_edge_check_batch.argtypes = [C.POINTER(_Obstacles), C.c_long, c_dp, c_dp,
                                           C.c_int, C.c_double, C.c_double, C.c_int,
                                           c_u8p, c_dp, c_dp, C.c_int]
        L.ref_point_check_batch.argtypes = [C.POINTER(_Obstacles), C.c_long, c_dp,
                                            C.c_double, C.c_double, C.c_int, c_u8p, C.c_int]
        L.ref_edge_check_timed_batch.argtypes = [C.POINTER(_TimedObstacles), C.c_long, c_dp, c_dp,
                                                 C.c_double, c_u8p, c_dp]
        L.ref_num_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _d(a):
    return a.ctypes.data_as(c_dp)


def _i(a):
    return a.ctypes.data_as(c_ip)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


def num_threads():
    return int(lib().ref_num_threads())


def metric(m, a, b):
    a = _f64(a); b = _f64(b)
    return float(lib().ref_metric(m, _d(a), _d(b)))


class RefTree:
    """Oracle twin of the reference KDTree (include/DRRT/kdtree.h:13-181)."""

    def __init__(self, dims=3, wraps=(2,), wrap_points=(2.0 * REF_PI,), metric=METRIC_R2S1):
        self.dims = dims
        self.metric = metric
        w = np.asarray(wraps, dtype=np.int32)
        wp = np.asarray(wrap_points, dtype=np.float64)
        self._h = lib().ref_tree_create(dims, len(w), _i(w), _d(wp), metric)
        if not self._h:
            raise ValueError("oracle: unsupported tree configuration")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().ref_tree_destroy(self._h)
            self._h = None

    @property
    def size(self):
        return lib().ref_tree_size(self._h)

    def insert(self, pos):
        pos = _f64(pos, (-1, self.dims))
        return lib().ref_tree_insert(self._h, pos.shape[0], _d(pos))

    def structure(self):
        n = self.size
        arrs = [np.empty(n, dtype=np.int32) for _ in range(4)]
        lib().ref_tree_structure(self._h, *[_i(a) for a in arrs])
        return dict(parent=arrs[0], left=arrs[1], right=arrs[2], split=arrs[3])

    def range_one(self, q, r, brute=False):
        """(ids, dists) in the oracle's discovery order (brute: sorted by id)."""
        q = _f64(q)
        cap = max(self.size, 1)
        ids = np.empty(cap, dtype=np.int32)
        dist = np.empty(cap, dtype=np.float64)
        f = lib().ref_range_brute if brute else lib().ref_range
        n = f(self._h, _d(q), float(r), _i(ids), _d(dist), cap)
        return ids[:n].copy(), dist[:n].copy()

    def last_visits(self):
        return int(lib().ref_range_last_visits())

    def range_batch(self, q, r, sort_by_id=True, nthreads=0):
        """CSR (offsets int64[nq+1], ids int32, dist f64), each row sorted by id."""
        q = _f64(q, (-1, self.dims))
        nq = q.shape[0]
        r = np.ascontiguousarray(np.broadcast_to(np.asarray(r, dtype=np.float64), (nq,)))
        counts = np.zeros(nq, dtype=np.int32)
        lib().ref_range_batch_count(self._h, nq, _d(q), _d(r), _i(counts), nthreads)
        off = np.zeros(nq + 1, dtype=np.int64)
        np.cumsum(counts, out=off[1:])
        ids = np.empty(int(off[-1]), dtype=np.int32)
        dist = np.empty(int(off[-1]), dtype=np.float64)
        lib().ref_range_batch_fill(self._h, nq, _d(q), _d(r), off.ctypes.data_as(c_llp),
                                   _i(ids), _d(dist), 1 if sort_by_id else 0, nthreads)
        return off, ids, dist

    def range_counts(self, q, r, nthreads=0):
        q = _f64(q, (-1, self.dims))
        nq = q.shape[0]
        r = np.ascontiguousarray(np.broadcast_to(np.asarray(r, dtype=np.float64), (nq,)))
        counts = np.zeros(nq, dtype=np.int32)
        lib().ref_range_batch_count(self._h, nq, _d(q), _d(r), _i(counts), nthreads)
        return counts

    def nearest_batch(self, q, nthreads=0):
        q = _f64(q, (-1, self.dims))
        nq = q.shape[0]
        ids = np.empty(nq, dtype=np.int32)
        dist = np.empty(nq, dtype=np.float64)
        lib().ref_nearest_batch(self._h, nq, _d(q), _i(ids), _d(dist), nthreads)
        return ids, dist

    def knn_batch(self, q, k, nthreads=0):
        q = _f64(q, (-1, self.dims))
        nq = q.shape[0]
        ids = np.empty((nq, k), dtype=np.int32)
        dist = np.empty((nq, k), dtype=np.float64)
        lib().ref_knn_batch(self._h, nq, _d(q), k, _i(ids), _d(dist), nthreads)
        return ids, dist

    def ghosts(self, q, best_dist):
        q = _f64(q)
        out = np.empty((7, self.dims), dtype=np.float64)
        n = lib().ref_ghosts(self._h, _d(q), float(best_dist), _d(out), 7)
        return out[:n].copy()


class RefObstacles:
    """Flat obstacle table; fields as the analytic checker reads them
    (O->kind_, obstacle_used_, life_span_, origin_, radius_, shape_.GetPolygon())."""

    def __init__(self, kind, used, life_span, origin, radius, vert_off, verts):
        self.kind = np.ascontiguousarray(kind, dtype=np.int32)
        self.used = np.ascontiguousarray(used, dtype=np.uint8)
        self.life_span = _f64(life_span)
        self.origin = _f64(origin, (-1, 2))
        self.radius = _f64(radius)
        self.vert_off = np.ascontiguousarray(vert_off, dtype=np.int32)
        self.verts = _f64(verts, (-1, 2))
        self.n = len(self.kind)
        self._s = _Obstacles(self.n, _i(self.kind), self.used.ctypes.data_as(c_u8p),
                             _d(self.life_span), _d(self.origin), _d(self.radius),
                             _i(self.vert_off), _d(self.verts))

    def edge_check_2d(self, o, p0, p1, radius):
        p0 = _f64(p0); p1 = _f64(p1)
        return bool(lib().ref_edge_check_2d(C.byref(self._s), o, _d(p0), _d(p1),
                                            float(radius), None))

    def edge_check_batch(self, start, end, edge_kind=0, robot_radius=0.5, r_min=1.0,
                         metric=METRIC_R2S1, want_clearance=False, nthreads=0):
        start = _f64(start, (-1, 3)); end = _f64(end, (-1, 3))
        n = start.shape[0]
        verdict = np.zeros(n, dtype=np.uint8)
        length = np.zeros(n, dtype=np.float64)
        clr = np.zeros(n, dtype=np.float64) if want_clearance else None
        lib().ref_edge_check_batch(C.byref(self._s), n, _d(start), _d(end), edge_kind,
                                   float(robot_radius), float(r_min), metric,
                                   verdict.ctypes.data_as(c_u8p), _d(length),
                                   _d(clr) if clr is not None else None, nthreads)
        return (verdict, length, clr) if want_clearance else (verdict, length)

    def point_check_batch(self, pts, robot_radius=0.5, collision_distance=0.1,
                          metric=METRIC_R2S1, nthreads=0):
        pts = _f64(pts, (-1, 3))
        verdict = np.zeros(pts.shape[0], dtype=np.uint8)
        lib().ref_point_check_batch(C.byref(self._s), pts.shape[0], _d(pts),
                                    float(robot_radius), float(collision_distance), metric,
                                    verdict.ctypes.data_as(c_u8p), nthreads)
        return verdict


class RefTimedObstacles:
    """kinds 6 / 7 of ExplicitEdgeCheck2D (obstacle.cpp:805-875): obstacles moving along a path of
    (dx, dy, time) rows relative to origin_; segments are (x, y, time) -> (x, y, time)"""

    def __init__(self, kind, used, life_span, origin, radius, path_off, path):
        self.kind = np.ascontiguousarray(kind, dtype=np.int32)
        self.used = np.ascontiguousarray(used, dtype=np.uint8)
        self.life_span = _f64(life_span)
        self.origin = _f64(origin, (-1, 2))
        self.radius = _f64(radius)
        self.path_off = np.ascontiguousarray(path_off, dtype=np.int32)
        self.path = _f64(path, (-1, 3))
        self._s = _TimedObstacles(len(self.kind), _i(self.kind), self.used.ctypes.data_as(c_u8p),
                                  _d(self.life_span), _d(self.origin), _d(self.radius), _i(self.path_off),
                                  _d(self.path))

    def edge_check_batch(self, start, end, radius=0.5, want_clearance=False):
        start = _f64(start, (-1, 3)); end = _f64(end, (-1, 3))
        n = start.shape[0]
        verdict = np.zeros(n, dtype=np.uint8)
        clr = np.zeros(n, dtype=np.float64) if want_clearance else None
        lib().ref_edge_check_timed_batch(C.byref(self._s), n, _d(start), _d(end), float(radius),
                                         verdict.ctypes.data_as(c_u8p), _d(clr) if clr is not None else None)
        return (verdict, clr) if want_clearance else verdict


def dubins_trajectory(start, end, r_min=1.0, metric=METRIC_R2S1):
    start = _f64(start); end = _f64(end)
    t = _Traj()
    lib().ref_dubins_trajectory(_d(start), _d(end), float(r_min), metric, C.byref(t))
    n = t.npts
    return dict(type=t.type.decode(), dist=t.dist,
                xy=np.stack([np.array(t.x[:n]), np.array(t.y[:n])], axis=1) if n else np.zeros((0, 2)))


def line_trajectory(n1, n2):
    n1 = _f64(n1); n2 = _f64(n2)
    t = _Traj()
    lib().ref_line_trajectory(_d(n1), _d(n2), C.byref(t))
    n = t.npts
    return np.stack([np.array(t.x[:n]), np.array(t.y[:n])], axis=1)


def dist_sqrd_point_segment(p, s, e):
    p = _f64(p); s = _f64(s); e = _f64(e)
    return float(lib().ref_dist_sqrd_point_segment(_d(p), _d(s), _d(e)))


def segment_dist_sqrd(pa, pb, qa, qb):
    pa = _f64(pa); pb = _f64(pb); qa = _f64(qa); qb = _f64(qb)
    return float(lib().ref_segment_dist_sqrd(_d(pa), _d(pb), _d(qa), _d(qb)))


def point_in_polygon(p, poly):
    p = _f64(p); poly = _f64(poly, (-1, 2))
    return bool(lib().ref_point_in_polygon(_d(p), _d(poly), poly.shape[0]))
