#!/usr/bin/env python
"""Measured FP64 peak of this GPU (DFMA, and the uncontracted DMUL + DADD the library issues).

    python benchmarks/fp64_peak.py [--device 0]      -> one JSON line

bench.py calls measure() and divides the edge-check kernels' FP64 work by it ("measure
one with a microbenchmark").  The number of one run is committed under profiles/.
"""
import ctypes as C
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libfp64peak.so")
SRC = os.path.join(HERE, "fp64_peak.cu")


def build(force=False):
    if force or not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
                               "-shared", "-Xcompiler", "-fPIC", "-cudart", "static", "-o", SO, SRC])
    return SO


def measure(device=0, reps=5):
    lib = C.CDLL(build())
    lib.fp64_peak_tflops.restype = C.c_double
    lib.fp64_peak_tflops.argtypes = [C.c_int, C.c_int, C.c_int]
    dfma = lib.fp64_peak_tflops(device, 0, reps)
    split = lib.fp64_peak_tflops(device, 1, reps)
    if dfma <= 0 or split <= 0:
        raise RuntimeError("fp64_peak: no usable CUDA device")
    return {"dfma_tflops": dfma, "dmul_dadd_tflops": split,
            "fp64_inst_per_s": dfma / 2.0 * 1e12,
            "how": "8 independent chains per thread, 8 x 256-thread CTAs per SM, best of %d launches, "
                   "CUDA events; DFMA = 2 flop per instruction, DMUL + DADD (no contraction, as the "
                   "library is built) = 1 flop per instruction" % reps}


if __name__ == "__main__":
    dev = int(sys.argv[sys.argv.index("--device") + 1]) if "--device" in sys.argv else 0
    print(json.dumps(measure(dev)))
