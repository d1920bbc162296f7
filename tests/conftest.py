import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def pt():
    """The product package (hyphenated name -> importlib)."""
    mod = importlib.import_module("performance-test_b200")
    exe = os.path.join(os.path.dirname(mod.HOST_LIB), "dolfinx-scaling-test")
    if not (os.path.exists(mod.HOST_LIB) and os.path.exists(mod.ABI_LIB) and os.path.exists(exe)):
        mod.build()
    return mod


@pytest.fixture(scope="session")
def oracle():
    import oracle as o
    o.build()
    return o


class NoBC:
    """View of a problem with the Dirichlet set removed (for the pre-BC invariants K4, K8)."""

    def __init__(self, P):
        self._P = P

    def __getattr__(self, k):
        return getattr(self._P, k)

    def __getitem__(self, k):
        import numpy as np
        return np.zeros(0, np.int32) if k == "bc_dofs" else self._P[k]


@pytest.fixture(scope="session")
def nobc():
    return NoBC
