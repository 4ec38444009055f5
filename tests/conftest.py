import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA GPU (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def oracle():
    from tests import oracle_ffi

    return oracle_ffi


@pytest.fixture(scope="session")
def ctx():
    import resvg_b200

    c = resvg_b200.Context(0)
    yield c
    c.close()
