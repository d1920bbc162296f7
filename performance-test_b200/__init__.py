"""performance-test_b200: B200-native hot path of FEniCS/performance-test.

Python is only the test/bench harness here: it binds the two in-tree shared libraries with ctypes.

* ``host``  -> ``libptb200_host.so`` (include/ptb200_host.h): stand-in for the DOLFINx setup the
  reference does on the host before its timed regions (mesh, dofmap, BCs, RHS, sparsity).
* ``abi``   -> ``libptb200.so`` (include/ptb200.h): the drop-in C-ABI over the sm_100a CUDA kernels
  that replace ``ZZZ Assemble matrix`` / ``ZZZ Assemble vector`` / ``ZZZ Solve``.

The package name contains a hyphen (it mirrors the reference repo's name), so import it with
``importlib.import_module("performance-test_b200")``.

There is no CPU fallback: anything that needs the CUDA library raises if it is missing.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB = os.path.join(_HERE, "libptb200_host.so")
ABI_LIB = os.path.join(_HERE, "libptb200.so")
INCLUDE_DIR = os.path.join(os.path.dirname(_HERE), "include")


def build(verbose: bool = False) -> None:
    """Compile both libraries in-tree (nvcc cross-compiles sm_100a without a GPU)."""
    r = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
        print(r.stderr)
    if r.returncode != 0:
        raise RuntimeError("performance-test_b200: build failed")


from . import host  # noqa: E402
from . import abi  # noqa: E402
from . import dist  # noqa: E402

__all__ = ["build", "host", "abi", "dist", "HOST_LIB", "ABI_LIB", "INCLUDE_DIR"]
