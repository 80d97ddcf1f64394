import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    return g.load_package()


@pytest.fixture(scope="session")
def ora():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def engine(pkg):
    """One context on cuda:0 through the C-ABI; fails loudly when the library or GPU is missing."""
    import __graft_entry__ as g
    so = os.path.join(g.PKG_DIR, "csrc", "libextfem_cuda.so")
    if not os.path.exists(so):
        g.build()
    eng = pkg.lib.Engine(0)
    yield eng
    eng.close()
