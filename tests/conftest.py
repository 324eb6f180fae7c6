import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ctx():
    """One accmsm context on cuda:0.  Fails loudly (no skip, no CPU path) if the extension or GPU is missing."""
    import accumulation_b200 as ab
    c = ab.Context(0)
    yield c
    c.close()
