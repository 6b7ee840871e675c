import ctypes
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_devices() -> int:
    """Number of CUDA devices the driver sees (0 without a driver); no torch import, no context."""
    try:
        cu = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        if cu.cuInit(0) != 0 or cu.cuDeviceGetCount(ctypes.byref(n)) != 0:
            return 0
        return n.value
    except OSError:
        return 0


def pytest_collection_modifyitems(config, items):
    # `pytest tests` on a machine without a GPU: the gpu tests are skipped, not failed.  The library itself still has
    # no CPU path (pk_create returns PK_E_NO_DEVICE; tests/test_abi_cpu.py).
    if _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
