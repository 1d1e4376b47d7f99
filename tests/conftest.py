import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def ctx():
    """One library context on cuda:0; fails loudly (no CPU fallback) when the extension or the GPU is missing."""
    import motionplanning_5d_m_b200 as M
    c = M.Context(0)
    yield c
    c.close()
