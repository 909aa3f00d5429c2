import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """The oracle is built on demand (gcc only); the CUDA library must already exist
    (__graft_entry__.build()) — importing rscm_b200 fails loudly otherwise."""
    from oracle import oracle

    oracle.build()
    yield
