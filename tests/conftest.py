import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built_lib():
    """The in-tree sm_100a library; built here if stale (nvcc cross-compiles without a GPU)."""
    from gdkvm_b200 import _build
    try:
        return _build.build()
    except Exception:
        if os.path.exists(_build.LIB_PATH):   # GPU box without write access / nvcc: use what travelled
            return _build.LIB_PATH
        raise


@pytest.fixture(scope="session")
def c_oracle():
    from oracle import c_oracle as co
    co.build()
    return co
