"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a,
loads, exports every symbol include/drrt_gpu.h declares, refuses to run
without a GPU (no CPU fallback), and the product never touches oracle/."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "drrt_gpu.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"DRRT_API[^;(]*?\b(drrt_\w+)\s*\(", src)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for need in ("drrt_store_create", "drrt_store_insert", "drrt_range_batch", "drrt_range_fetch",
                 "drrt_nearest_batch", "drrt_knn_batch", "drrt_obstacles_set",
                 "drrt_edgecheck_batch", "drrt_edgecheck_ids_batch", "drrt_pointcheck_batch",
                 "drrt_dubins_trajectory_batch"):
        assert need in syms
    # every entry point cites the reference interface it replaces (file:line)
    src = open(HEADER).read()
    assert len(re.findall(r"(kdtree|obstacle|dubinsedge|ghostpoint|smalltest|thetastar)\.(cpp|h):\d+", src)) >= 15


def test_library_exports_every_declared_symbol(drrt):
    lib = ctypes.CDLL(drrt.capi.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    # and the ctypes binding covers the same set
    assert sorted(drrt.capi.SIGNATURES) == declared_symbols()
    assert lib.drrt_abi_version() == drrt.capi.ABI_VERSION == 4


def test_library_is_sm100a_cuda_code(drrt):
    out = subprocess.run(["cuobjdump", "-lelf", drrt.capi.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_no_cpu_fallback(drrt):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(drrt.DrrtError) as ei:
        drrt.KDTree()
    assert ei.value.code == drrt.capi.ERR_NO_DEVICE
    assert "no CPU path" in str(ei.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "drrt_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "liboracle" not in txt and "oracle/" not in txt, f
    assert "oracle" not in open(HEADER).read()
