import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "one_leg_golden.npz"))


@pytest.fixture(scope="session")
def port():
    from oracle.oracle import PortOracle
    return PortOracle()


@pytest.fixture(scope="session")
def ref():
    from oracle.oracle import RefOracle, have_ref
    if not have_ref():
        pytest.skip("oracle/_ref/libref_oracle.so not built (needs /root/reference)")
    return RefOracle()


@pytest.fixture(scope="session")
def oracle():
    """Strongest oracle available: compiled reference if shipped, else the pinned C port."""
    from oracle.oracle import best
    return best()


@pytest.fixture(scope="session")
def lrm():
    import lrm_loader
    lrm_loader.build()
    return lrm_loader.load()
