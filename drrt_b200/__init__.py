"""drrt_b200 -- B200-native (sm_100a) back end for the data-parallel hot path of
cosa0481/DRRT: batched KD-tree neighbourhood queries and edge collision checks.

The product is the C-ABI library (include/drrt_gpu.h, built from csrc/ by
`python -m drrt_b200.build`); this package is the thin host-side mirror of the
reference's KDTree / edge-check interface used by the tests and the benchmark.
"""
from . import capi
from .capi import (EDGE_DUBINS, EDGE_STRAIGHT, INF, METRIC_EUCLID, METRIC_R2S1_WRAP, PI,
                   DrrtError)
from .kdtree import KDTree

__all__ = ["KDTree", "capi", "DrrtError", "METRIC_R2S1_WRAP", "METRIC_EUCLID", "EDGE_DUBINS",
           "EDGE_STRAIGHT", "PI", "INF"]
