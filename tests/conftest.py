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


def load_golden(name):
    """(Problem parsed from the committed .in, dict of reference dumps)."""
    from oofem_b200.inputfile import read_input
    pb = read_input(os.path.join(GOLDEN, name + ".in"))
    d = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    return pb, d


@pytest.fixture(scope="session")
def golden():
    return load_golden


def relerr(a, b):
    """max |a-b| / max |b|  (the norm-wise 'relative' used for matrices and vectors)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.abs(b).max() if b.size else 0.0
    if den == 0.0:
        return float(np.abs(a).max()) if a.size else 0.0
    return float(np.abs(a - b).max() / den)
