import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as ge  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    """The package, with libreo_cuda.so built if this checkout has not been built yet (nvcc cross-compiles)."""
    so = os.path.join(ge.PKG_DIR, "libreo_cuda.so")
    if not os.path.exists(so):
        ge.build()
    return ge.load_package()


@pytest.fixture(scope="session")
def oracle():
    """numpy restatement (oracle/reo_oracle.py) -- checker only."""
    return ge.load_oracle()[0]


@pytest.fixture(scope="session")
def coracle():
    """plain-C restatement (oracle/reo_oracle.c via ctypes) -- checker only."""
    return ge.load_oracle()[1]


@pytest.fixture(scope="session")
def reo(pkg):
    """One libreo_cuda handle on cuda:0, tie seed 7."""
    h = pkg.Reo(0, seed=pkg.synth.TIE_SEED)
    yield h
    h.close()


def small_case(seed, r, n1, n2, scale=8, n3=0):
    """Count data with many ties; optional third group."""
    rng = np.random.default_rng(seed)
    mu = np.exp(rng.normal(2.0, 1.5, r))
    c = n1 + n2 + n3
    lam = mu[:, None] * np.ones((1, c))
    de = rng.choice(r, max(r // 10, 1), replace=False)
    lam[de[: len(de) // 2], n1:n1 + n2] *= 3.0
    lam[de[len(de) // 2:], :n1] *= 3.0
    data = (rng.poisson(lam) // scale).astype(np.int64)
    group = ["a"] * n1 + ["b"] * n2 + ["c"] * n3
    perm = rng.permutation(c)  # interleave the groups: staging must sort samples by level
    return data[:, perm], [group[i] for i in perm]


GOLDEN = os.path.join(ROOT, "tests", "golden")
