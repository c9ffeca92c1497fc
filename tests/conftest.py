import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def prv():
    """The product package (ctypes binding over libprv_b200.so).  Builds the library if nvcc is here."""
    import load_pkg
    mod = load_pkg.load()
    if not os.path.exists(mod.LIB_PATH):
        mod.build()
    return mod


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    import oracle
    oracle.build()
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def synth(prv):
    from nerf_prv_b200 import synth as s
    return s


@pytest.fixture(scope="session")
def ctx(prv):
    c = prv.Context(0)
    yield c
    c.close()
