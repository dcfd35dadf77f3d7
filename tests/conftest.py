import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_usable() -> bool:
    try:
        import torch

        return bool(torch.cuda.is_available())
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """Plain `pytest tests` on a CPU-only box: gpu-marked tests are skipped, not errors."""
    if _cuda_usable():
        return
    skip = pytest.mark.skip(reason="no usable CUDA device (gpu-marked test)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def engine_built():
    """Build (if stale) and return the engine .so path; CPU-only containers can do this."""
    from envpool_b200 import _build

    return _build.build_all()


@pytest.fixture(scope="module")
def capi(engine_built):
    """The ctypes binding of the C ABI on a box with a GPU (every GPU parity test goes
    through it)."""
    import torch

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from envpool_b200 import _capi

    _capi.load_library()
    return _capi
