import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def rvio():
    import rvio_b200  # noqa: F401  (import shim -> r-vio_b200/)
    return rvio_b200


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """Tests marked `gpu` need a B200: on a machine without one they are skipped (not failed) unless `-m gpu` asked for
    them explicitly, in which case the product's own loud failure (RVIO_ERR_CUDA, no CPU fallback) is what should show."""
    if have_gpu() or "gpu" in (config.getoption("-m") or ""):
        return
    skip = pytest.mark.skip(reason="no CUDA device (B200) on this machine")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
