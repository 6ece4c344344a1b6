import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (loads libstacker_cuda.so through ctypes; fails if it is not built)."""
    import __graft_entry__ as ge
    so = os.path.join(ge.PKG_DIR, "libstacker_cuda.so")
    if not os.path.exists(so):
        ge.build()
    return ge.load_package()


@pytest.fixture(scope="session")
def have_cv2():
    try:
        import cv2  # noqa: F401
        return True
    except Exception:
        return False
