import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "icp-symm_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cat():
    return np.load(os.path.join(GOLD, "cat_xyz.npy"))


@pytest.fixture(scope="session")
def cat_out():
    return np.load(os.path.join(GOLD, "cat_out_xyz.npy"))


@pytest.fixture(scope="session")
def bunny():
    return np.load(os.path.join(GOLD, "bunny_xyz.npy"))


@pytest.fixture(scope="session")
def gold():
    return dict(np.load(os.path.join(GOLD, "cat_golden.npz")))


@pytest.fixture(scope="session")
def solve_gold():
    """independent fp64 numpy.linalg.lstsq solve of the canonical iteration on cat (tests/golden/make_golden.py)"""
    return dict(np.load(os.path.join(GOLD, "cat_solve_golden.npz")))
