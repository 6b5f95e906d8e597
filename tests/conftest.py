import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def po():
    """the CPU oracle (test infrastructure; never imported by drrt_b200)"""
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def drrt():
    """the product package; building it here keeps stale .so files out of the run"""
    from drrt_b200 import build
    build.build()
    import drrt_b200
    drrt_b200.capi.load()
    return drrt_b200


def csr_rows(off, ids):
    return [ids[off[i]:off[i + 1]] for i in range(len(off) - 1)]
