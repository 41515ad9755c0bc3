import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "np-sims_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def kat_top10():
    return np.load(os.path.join(GOLDEN, "kat_top10.npz"))


@pytest.fixture(scope="session")
def kat_hamming():
    return np.load(os.path.join(GOLDEN, "kat_hamming.npz"))


@pytest.fixture(scope="session")
def c1_glove():
    return np.load(os.path.join(GOLDEN, "c1_glove.npz"))


@pytest.fixture(scope="session")
def random_ties():
    return np.load(os.path.join(GOLDEN, "random_ties.npz"))
