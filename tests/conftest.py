import ctypes
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "multigpu: needs at least two CUDA devices (torchrun / NCCL)")


def cuda_device_count() -> int:
    """CUDA devices visible to this process, asked of the driver directly (no torch import, no context)."""
    try:
        cu = ctypes.CDLL("libcuda.so.1")
        if cu.cuInit(0) != 0:
            return 0
        n = ctypes.c_int(0)
        return n.value if cu.cuDeviceGetCount(ctypes.byref(n)) == 0 else 0
    except OSError:
        return 0


def pytest_collection_modifyitems(config, items):
    """``-m gpu`` tests are SKIPPED (not failed) on a box without a CUDA device, and the multi-GPU ones on a box
    with fewer than two; a plain ``pytest tests`` therefore passes on a CPU box.  With a device present nothing is
    skipped: a missing or broken extension then fails loudly inside the test."""
    ndev = cuda_device_count()
    no_gpu = pytest.mark.skip(reason="no CUDA device visible (the PMM engine has no CPU path)")
    few = pytest.mark.skip(reason="needs >= 2 CUDA devices, %d visible" % ndev)
    for it in items:
        if "gpu" in it.keywords and ndev == 0:
            it.add_marker(no_gpu)
        elif "multigpu" in it.keywords and ndev < 2:
            it.add_marker(few)
