import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_calls():
    from oracle import backends
    return backends.oracle()


def _ref(dim):
    from oracle import backends
    r = backends.reference(dim)
    if r is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    return r


@pytest.fixture(scope="session")
def ref3d():
    return _ref(3)


@pytest.fixture(scope="session")
def ref2d():
    return _ref(2)
