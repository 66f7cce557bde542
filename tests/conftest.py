import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def native_lib():
    """The built C-ABI library; building is part of the CPU suite (nvcc cross-compiles without a GPU)."""
    from regularizepsf_b200.csrc import build as native_build

    native_build.build()
    from regularizepsf_b200 import _native

    return _native.load()
