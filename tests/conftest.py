import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def po():
    """The CPU oracle (test infrastructure; see oracle/pixo_oracle.h)."""
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def lib():
    from pixo_b200 import build as _b
    _b.build()
    from pixo_b200 import _lib
    return _lib.load()


@pytest.fixture(scope="session")
def gpu_ctx(lib):
    import pixo_b200
    if lib.pixo_b200_device_count() < 1:
        pytest.fail("GPU test selected but no CUDA device is visible (no CPU fallback exists)")
    return pixo_b200.Context(0)
