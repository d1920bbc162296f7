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


class Perturbed:
    """View of a problem whose vertices are moved by up to `amp` cell widths (same topology, same
    Dirichlet set and source data): on the regular lattice many products of the element kernels
    vanish or coincide; a jittered mesh exercises every term, the orientation handling of the star
    walks included. Both sides of a parity test read the geometry through this view."""

    def __init__(self, P, amp=0.15, seed=11):
        import numpy as np
        self._P = P
        x = np.array(P["x"]).reshape(-1, 3)
        h = np.array([1.0 / P.nx, 1.0 / P.ny, 1.0 / P.nz])
        rng = np.random.default_rng(seed)
        # the jitter must be a function of the vertex position (ranks of a partition share vertices)
        key = np.round(x * np.array([P.nx, P.ny, P.nz])).astype(np.int64)
        gid = (key[:, 2] * (P.ny + 1) + key[:, 1]) * (P.nx + 1) + key[:, 0]
        table = rng.uniform(-amp, amp, size=((P.nx + 1) * (P.ny + 1) * (P.nz + 1), 3))
        self._x = np.ascontiguousarray((x + table[gid] * h).reshape(-1))
        # coordinates by dof for the vertex dofs (local dofs 0..3 of a Lagrange cell sit on its vertices)
        dm = np.array(P["dofmap"]).reshape(-1, P.nd)[:, :4]
        xd = np.array(P["x_dofmap"]).reshape(-1, 4)
        dof_x = np.array(P["dof_x"]).reshape(-1, 3).copy()
        dof_x[dm.reshape(-1)] = self._x.reshape(-1, 3)[xd.reshape(-1)]
        self._dof_x = np.ascontiguousarray(dof_x.reshape(-1))

    def __getattr__(self, k):
        return getattr(self._P, k)

    def __getitem__(self, k):
        if k == "x":
            return self._x
        if k == "dof_x":
            return self._dof_x
        return self._P[k]


@pytest.fixture(scope="session")
def perturbed():
    return Perturbed
