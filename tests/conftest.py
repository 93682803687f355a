"""Shared fixtures.  `-m "not gpu"` runs here on CPU; `-m gpu` runs on a B200."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


@pytest.fixture(scope="session")
def oracle():
    from oracle_py import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref_strict():
    """The compiled reference (oracle/_ref); only where it was prebuilt."""
    from oracle_py import Ref
    if not Ref.available("strict"):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return Ref("strict")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def rect_map(nx, ny, nobs, seed, lo=3, hi=9):
    """Same deterministic test maps as oracle/gen_golden.py."""
    g = np.random.default_rng(seed)
    occ = np.ones((ny, nx))
    for _ in range(nobs):
        x = int(g.integers(0, nx)); y = int(g.integers(0, ny))
        w = int(g.integers(lo, hi + 1)); h = int(g.integers(lo, hi + 1))
        occ[y:y + h, x:x + w] = 0
    return occ
