import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"  # absent on the GPU box: only `not gpu` tests may touch it, and they skip without it


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def lib():
    """The C-ABI library, built in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    from norma_b200 import build, ffi

    build.build()
    return ffi.load_library()


def golden(name):
    import numpy as np

    return np.load(os.path.join(GOLDEN, name))
