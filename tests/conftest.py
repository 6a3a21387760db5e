import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """Vectors produced by the unmodified reference (tests/golden/make_golden.py)."""
    return np.load(os.path.join(ROOT, "tests", "golden", "ref_golden.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle -- the checker, never the thing under test on GPU runs."""
    from oracle import f16_oracle

    f16_oracle.build()
    return f16_oracle
