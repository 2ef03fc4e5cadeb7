import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def native_lib():
    """libmbseg.so, built in-tree if the sources are newer (nvcc cross-compiles without a GPU)."""
    from microbeseg_b200 import build
    build.build_lib()
    from microbeseg_b200 import _native
    return _native.lib()
