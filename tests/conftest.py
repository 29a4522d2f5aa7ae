import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def cocg():
    import cocg as _cocg
    return _cocg


@pytest.fixture(scope="session")
def bn(cocg):
    ctx = cocg.Context(cocg.BN254, 0)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def bls(cocg):
    ctx = cocg.Context(cocg.BLS12_381, 0)
    yield ctx
    ctx.close()
