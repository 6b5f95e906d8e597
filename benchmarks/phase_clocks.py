#!/usr/bin/env python
"""Where a query's time goes inside range_cta_kernel: cycles between the phase marks of thread 0,
summed over all queries of config 2a (needs the experiment build

    python -m drrt_b200.build --variant phases -DDRRT_PHASE_CLOCKS
    DRRT_B200_LIB=$PWD/drrt_b200/variants/phases/libdrrt_b200.so python benchmarks/phase_clocks.py
)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import drrt_b200  # noqa: E402
from drrt_b200 import capi, synth  # noqa: E402

NAMES = [  # noqa: E501
         "clear + probe copy, barrier", "column runs, first scan, barrier", "run list, barrier",
         "thread 0's trips", "wait for the slowest warp", "tail scan, barrier", "bucket prefix (2 barriers)",
         "positions, barrier", "insertion sort, barrier", "thread 0's row write", "end barrier"]
dev = torch.device("cuda", 0)
lib = capi.load()
raw = C.CDLL(os.environ["DRRT_B200_LIB"])
raw.drrt_debug_phases.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
nodes, q = synth.nodes(), synth.queries()
tree = drrt_b200.KDTree(device=0, capacity=len(nodes))
tree.KDInsert(nodes)
qd = torch.from_numpy(q).to(dev)
r = synth.shrinking_radius(len(nodes), synth.GAMMA_REF)
out = (C.c_ulonglong * 16)()
for _ in range(3):
    tree.KDFindWithinRange_dev(r, qd)
torch.cuda.synchronize()
raw.drrt_debug_phases(out, 1)
steps = 5
for _ in range(steps):
    tree.KDFindWithinRange_dev(r, qd)
torch.cuda.synchronize()
raw.drrt_debug_phases(out, 0)
v = np.array(list(out)[:len(NAMES)], dtype=np.float64) / (steps * len(q))
print("cycles per query (thread 0's clock), total %.0f" % v.sum())
for n, x in zip(NAMES, v):
    print("  %-36s %8.0f  %5.1f %%" % (n, x, 100 * x / v.sum()))
