import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import ctypes
        rt = ctypes.CDLL("libcudart.so")
    except OSError:
        try:
            import torch
            return torch.cuda.is_available()
        except Exception:
            return False
    n = ctypes.c_int(0)
    return rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 and n.value > 0


HAS_GPU = None


def pytest_collection_modifyitems(config, items):
    global HAS_GPU
    if HAS_GPU is None:
        try:
            import torch
            HAS_GPU = torch.cuda.is_available()
        except Exception:
            HAS_GPU = _has_gpu()
    if HAS_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import binding
    binding.build()
    return binding


@pytest.fixture(scope="session")
def msl():
    from manhattanslam_b200 import build
    build.build()
    import manhattanslam_b200
    return manhattanslam_b200
