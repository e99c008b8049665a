import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_sessionstart(session):
    """Build what the tests load if it is not there yet (a fresh checkout has no .so: they are git-ignored).  nvcc cross-
    compiles sm_100a without a GPU; on the GPU box the prebuilt files travel with the snapshot and nothing is rebuilt.
    Building the oracle is not using it: only tests/, smoke() and bench.py's CPU legs call into it."""
    from sawfish_b200 import _lib
    from oracle import oracle as O
    if not os.path.exists(_lib.SO_PATH):
        _lib.build()
    O.build()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def ctx():
    """One AlignContext on cuda:0 for the whole GPU session. Fails loudly if the CUDA library is missing."""
    from sawfish_b200.batch import AlignContext
    c = AlignContext(0)
    yield c
    c.close()
