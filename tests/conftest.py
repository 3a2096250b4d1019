import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lib():
    """libcosma_b200.so, built on demand (nvcc cross-compiles without a GPU)."""
    from cosma_b200 import _lib
    return _lib.load(build=True)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.lib()
    return o


@pytest.fixture(scope="session")
def ref(oracle):
    if not oracle.have_ref():
        if os.path.exists("/root/reference/src/cosma/multiply.cpp"):
            oracle.build(ref=True)
        else:
            pytest.skip("oracle/_ref not built and /root/reference not present")
    oracle.ref()
    return oracle
