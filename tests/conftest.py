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
    """The CPU checker (test infrastructure): builds liboracle.so (and oracle/_ref when /root/reference exists)."""
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def crd():
    """The product package with the library built."""
    from crdmodel_b200 import build as B
    B.build()
    import crdmodel_b200
    crdmodel_b200.lib()
    return crdmodel_b200


@pytest.fixture(scope="session")
def ctx(crd):
    c = crd.Context(0)
    yield c
    c.close()
