"""ctypes binding of libdrrt_b200.so (include/drrt_gpu.h).

The shared object is built in-tree by drrt_b200.build (nvcc, sm_100a).  There is
no CPU fallback: loading fails loudly when the library is missing, and every
entry point fails with DRRT_ERR_NO_DEVICE when no sm_100 GPU is present.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DRRT_B200_LIB: another build of the same library (kernel A/B runs); still no CPU path
LIB_PATH = os.environ.get("DRRT_B200_LIB") or os.path.join(_HERE, "libdrrt_b200.so")

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NOMEM, ERR_NO_DEVICE, ERR_STATE, ERR_CAPACITY = range(8)
METRIC_R2S1_WRAP, METRIC_EUCLID = 0, 1
EDGE_DUBINS, EDGE_STRAIGHT = 0, 1
TRAJ_MAX = 200
ABI_VERSION = 4
PI = 3.1415926536  # include/DRRT/distancefunctions.h:22
INF = 1000000000000.0

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_u8p = C.POINTER(C.c_uint8)
c_i64p = C.POINTER(C.c_int64)
c_fp = C.POINTER(C.c_float)


class RangeResult(C.Structure):
    _fields_ = [("nq", C.c_int64), ("total", C.c_int64), ("offsets_dev", C.c_void_p),
                ("ids_dev", C.c_void_p), ("dist_dev", C.c_void_p), ("counts_dev", C.c_void_p),
                ("extent", C.c_int64), ("packed", C.c_int32)]


class DrrtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("drrt_b200 error %d: %s" % (code, msg))
        self.code = code


# name -> (restype, argtypes); mirrors include/drrt_gpu.h one to one
SIGNATURES = {
    "drrt_last_error": (C.c_char_p, []),
    "drrt_abi_version": (C.c_int, []),
    "drrt_kernel_launch_count": (C.c_int64, []),
    "drrt_store_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), c_dp, C.c_int,
                                    C.c_int64, C.POINTER(C.c_void_p)]),
    "drrt_store_destroy": (C.c_int, [C.c_void_p]),
    "drrt_store_size": (C.c_int, [C.c_void_p, c_i64p]),
    "drrt_store_stream": (C.c_void_p, [C.c_void_p]),
    "drrt_store_insert": (C.c_int, [C.c_void_p, C.c_int64, c_dp, c_ip]),
    "drrt_store_insert_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, c_ip]),
    "drrt_store_structure": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, c_ip, c_ip, c_ip, c_ip]),
    "drrt_store_positions": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, c_dp]),
    "drrt_range_batch": (C.c_int, [C.c_void_p, C.c_int64, c_dp, c_dp, C.c_double, c_ip, c_i64p]),
    "drrt_range_fetch": (C.c_int, [C.c_void_p, c_ip, c_dp]),
    "drrt_range_batch_into": (C.c_int, [C.c_void_p, C.c_int64, c_dp, c_dp, C.c_double, C.c_int64,
                                        c_i64p, c_ip, c_dp, c_i64p]),
    "drrt_range_batch_into_f32": (C.c_int, [C.c_void_p, C.c_int64, c_dp, c_dp, C.c_double, C.c_int64,
                                            c_i64p, c_ip, C.POINTER(C.c_float), c_i64p]),
    "drrt_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "drrt_host_free": (C.c_int, [C.c_void_p]),
    "drrt_lmc_write": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, c_dp]),
    "drrt_lmc_scatter": (C.c_int, [C.c_void_p, C.c_int64, c_ip, c_dp]),
    "drrt_lmc_read": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, c_dp]),
    "drrt_comm_init": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(C.c_int), c_dp,
                                 C.c_int, C.c_int64, C.POINTER(C.c_void_p)]),
    "drrt_comm_destroy": (C.c_int, [C.c_void_p]),
    "drrt_comm_size": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "drrt_comm_store": (C.c_void_p, [C.c_void_p, C.c_int]),
    "drrt_comm_broadcast_bytes": (C.c_int64, [C.c_void_p]),
    "drrt_comm_insert": (C.c_int, [C.c_void_p, C.c_int64, c_dp, c_ip]),
    "drrt_comm_obstacles_set": (C.c_int, [C.c_void_p, C.c_int32, c_ip, c_u8p, c_dp, c_dp, c_dp, c_ip,
                                          c_dp]),
    "drrt_comm_range_batch_into": (C.c_int, [C.c_void_p, C.c_int64, c_dp, c_dp, C.c_double, C.c_int64,
                                             c_i64p, c_ip, c_dp, c_i64p]),
    "drrt_comm_range_batch_into_f32": (C.c_int, [C.c_void_p, C.c_int64, c_dp, c_dp, C.c_double,
                                                 C.c_int64, c_i64p, c_ip, C.POINTER(C.c_float), c_i64p]),
    "drrt_comm_nearest_batch": (C.c_int, [C.c_void_p, C.c_int64, c_dp, c_ip, c_dp]),
    "drrt_comm_knn_batch": (C.c_int, [C.c_void_p, C.c_int64, c_dp, C.c_int, c_ip, c_dp]),
    "drrt_comm_edgecheck_batch": (C.c_int, [C.c_void_p, C.c_int64, c_dp, c_dp, C.c_int, C.c_double,
                                            C.c_double, C.c_int, c_u8p, c_dp]),
    "drrt_comm_edgecheck_ids_batch": (C.c_int, [C.c_void_p, C.c_int64, c_ip, c_ip, C.c_int, C.c_double,
                                                C.c_double, C.c_int, c_u8p, c_dp]),
    "drrt_comm_pointcheck_batch": (C.c_int, [C.c_void_p, C.c_int64, c_dp, C.c_double, C.c_double,
                                             c_u8p]),
    "drrt_comm_lmc_write": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, c_dp]),
    "drrt_comm_lmc_scatter": (C.c_int, [C.c_void_p, C.c_int64, c_ip, c_dp]),
    "drrt_comm_extend_batch": (C.c_int, [C.c_void_p, C.c_int64, c_dp, c_dp, C.c_double, C.c_double,
                                         C.c_double, C.c_int, c_dp, c_ip, c_i64p, c_ip, c_dp]),
    "drrt_range_batch_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_double,
                                       C.c_void_p, C.POINTER(RangeResult)]),
    "drrt_range_pack": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(RangeResult)]),
    "drrt_nearest_batch": (C.c_int, [C.c_void_p, C.c_int64, c_dp, c_ip, c_dp]),
    "drrt_nearest_batch_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p]),
    "drrt_knn_batch": (C.c_int, [C.c_void_p, C.c_int64, c_dp, C.c_int, c_ip, c_dp]),
    "drrt_knn_batch_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "drrt_obstacles_set": (C.c_int, [C.c_void_p, C.c_int32, c_ip, c_u8p, c_dp, c_dp, c_dp, c_ip,
                                     c_dp]),
    "drrt_edgecheck_batch": (C.c_int, [C.c_void_p, C.c_int64, c_dp, c_dp, C.c_int, C.c_double,
                                       C.c_double, C.c_int, c_u8p, c_dp]),
    "drrt_edgecheck_ids_batch": (C.c_int, [C.c_void_p, C.c_int64, c_ip, c_ip, C.c_int, C.c_double,
                                           C.c_double, C.c_int, c_u8p, c_dp]),
    "drrt_edgecheck_batch_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                           C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "drrt_edgecheck_subset_batch": (C.c_int, [C.c_void_p, C.c_int64, c_dp, c_dp, c_ip, c_ip, C.c_int,
                                              C.c_double, C.c_double, c_u8p, C.c_int32, c_u8p, c_dp]),
    "drrt_edgecheck_subset_batch_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                                  C.c_void_p, C.c_void_p, C.c_int, C.c_double,
                                                  C.c_double, c_u8p, C.c_int32, C.c_void_p,
                                                  C.c_void_p, C.c_void_p]),
    "drrt_timed_obstacles_set": (C.c_int, [C.c_void_p, C.c_int32, c_ip, c_u8p, c_dp, c_dp, c_dp, c_ip, c_dp]),
    "drrt_edgecheck2d_timed_batch": (C.c_int, [C.c_void_p, C.c_int64, c_dp, c_dp, C.c_double, c_u8p]),
    "drrt_edgecheck2d_timed_batch_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_double,
                                                   C.c_void_p, C.c_void_p]),
    "drrt_obstacles_move": (C.c_int, [C.c_void_p, C.c_int32, c_ip, c_dp, C.c_int]),
    "drrt_obstacles_update": (C.c_int, [C.c_void_p, C.c_int32, c_ip, c_u8p, c_dp]),
    "drrt_edges_append": (C.c_int, [C.c_void_p, C.c_int64, c_ip, c_ip, c_i64p]),
    "drrt_edges_count": (C.c_int, [C.c_void_p, c_i64p]),
    "drrt_edges_clear": (C.c_int, [C.c_void_p]),
    "drrt_obstacle_sweep": (C.c_int, [C.c_void_p, C.c_int32, c_u8p, C.c_int32, C.c_int, C.c_double,
                                      C.c_double, c_i64p, c_i64p, c_i64p]),
    "drrt_obstacle_sweep_fetch": (C.c_int, [C.c_void_p, c_i64p, c_u8p]),
    "drrt_obstacle_conflict_batch": (C.c_int, [C.c_void_p, C.c_int32, c_ip, C.c_double, c_ip, c_i64p]),
    "drrt_extend_batch": (C.c_int, [C.c_void_p, C.c_int64, c_dp, c_dp, C.c_double, C.c_double,
                                    C.c_double, C.c_int, c_dp, c_ip, c_i64p, c_ip, c_dp]),
    "drrt_extend_fetch": (C.c_int, [C.c_void_p, c_ip, c_dp, c_u8p, c_dp]),
    "drrt_extend_rewire": (C.c_int, [C.c_void_p, C.c_double, c_i64p]),
    "drrt_extend_rewire_fetch": (C.c_int, [C.c_void_p, c_ip, c_ip, c_dp]),
    "drrt_extend_batch_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_double,
                                        C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                        C.POINTER(RangeResult), C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p]),
    "drrt_dubins_trajectory_batch": (C.c_int, [C.c_void_p, C.c_int64, c_dp, c_dp, C.c_double,
                                               c_dp, c_ip, C.c_char_p, c_dp]),
    "drrt_pointcheck_batch": (C.c_int, [C.c_void_p, C.c_int64, c_dp, C.c_double, C.c_double,
                                        c_u8p]),
    "drrt_pointcheck_batch_dev": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_double,
                                            C.c_double, C.c_void_p, C.c_void_p]),
}

_lib = None


def load():
    """Loads the C-ABI library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "drrt_b200: %s is missing -- run `python -m drrt_b200.build` "
                "(needs nvcc; there is no CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        if lib.drrt_abi_version() != ABI_VERSION:
            raise ImportError("drrt_b200: ABI version mismatch")
        _lib = lib
    return _lib


def check(rc):
    if rc != OK:
        raise DrrtError(rc, load().drrt_last_error().decode(errors="replace"))


def launches():
    return int(load().drrt_kernel_launch_count())
