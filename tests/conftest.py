import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "leg-kilo_b200", "python"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import ctypes
        rt = ctypes.CDLL("libcudart.so")
        n = ctypes.c_int(0)
        return rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        try:
            import torch
            return torch.cuda.is_available()
        except Exception:
            return False


HAS_GPU = None


def pytest_collection_modifyitems(config, items):
    global HAS_GPU
    if HAS_GPU is None:
        HAS_GPU = _has_gpu()
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device here")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
