import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return {name[:-4]: np.load(os.path.join(GOLDEN, name)) for name in os.listdir(GOLDEN) if name.endswith(".npz")}


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Both the product library and the CPU checkers must exist before any test runs."""
    import __graft_entry__ as g

    g.build()
