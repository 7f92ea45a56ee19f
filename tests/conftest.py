import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """the CPU oracle (test infrastructure only)"""
    import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def g():
    import gmres_b200
    return gmres_b200


@pytest.fixture(scope="session")
def ctx(g):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    c = g.Context(0)
    yield c
    c.close()
