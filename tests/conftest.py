import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device here")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _build_checkers():
    """The CPU checkers (oracle/) are test infrastructure: build them once per session."""
    from oracle import oracle as orc
    orc.build()
    yield


@pytest.fixture(scope="session")
def libekf():
    """libekfcuda.so must exist (built in-tree by __graft_entry__.build / slam_ros_b200.build)."""
    from slam_ros_b200 import build as b
    import shutil
    if shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"):
        b.build()
    from slam_ros_b200 import load_library
    return load_library()
