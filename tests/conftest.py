import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def lz():
    import __graft_entry__ as g
    if not os.path.exists(os.path.join(g.PKG_DIR, "csrc", "liblzb200.so")):
        g.build()
    return g.load_package()


@pytest.fixture(scope="session")
def orc():
    import oracle
    return oracle


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

    def load(name):
        return np.load(os.path.join(d, name + ".npz"))
    return load
