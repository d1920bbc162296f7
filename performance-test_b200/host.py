"""ctypes binding of libptb200_host.so (include/ptb200_host.h)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_lib = None
_DT = {0: np.float64, 1: np.int32, 2: np.int64}

SCALARS = ["n_cells", "n_cells_owned", "n_ghost_cells_front", "cell_global_offset",
           "n_cells_global", "n_vertices", "nd", "bs", "order", "n_owned", "n_ghost", "n_global",
           "global_offset", "nnz", "n_bc", "n_facets", "n_nbr", "rank", "nranks", "nx", "ny", "nz"]
ARRAYS = ["x", "x_dofmap", "dofmap", "dof_x", "ghost_global", "ghost_owner", "rowptr", "cols",
          "bc_dofs", "f", "g", "facet_cells", "facet_local", "nbr_ranks", "send_displ",
          "recv_displ", "local_indices", "remote_indices"]


def lib():
    global _lib
    if _lib is None:
        from . import HOST_LIB
        if not os.path.exists(HOST_LIB):
            raise RuntimeError(f"{HOST_LIB} is missing: run __graft_entry__.build() / make")
        L = C.CDLL(HOST_LIB)
        L.pth_last_error.restype = C.c_char_p
        L.pth_num_entities.argtypes = [C.c_int64] * 3 + [C.c_int, C.POINTER(C.c_int64)]
        L.pth_num_pdofs.argtypes = [C.c_int64] * 3 + [C.c_int, C.c_int, C.POINTER(C.c_int64)]
        L.pth_cube_sizing.argtypes = [C.c_uint64, C.c_int, C.c_uint64, C.c_int, C.c_uint64,
                                      C.POINTER(C.c_int64)]
        L.pth_problem_create.argtypes = [C.c_char_p, C.c_int, C.c_int64, C.c_int64, C.c_int64,
                                         C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.pth_problem_create_sizes_only.argtypes = L.pth_problem_create.argtypes
        L.pth_problem_renumber.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
        L.pth_problem_destroy.argtypes = [C.c_void_p]
        L.pth_problem_destroy.restype = None
        L.pth_problem_scalar.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64)]
        L.pth_problem_array.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_int64), C.POINTER(C.c_int)]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise RuntimeError(lib().pth_last_error().decode())


def num_entities(i, j, k, nrefine=0):
    out = (C.c_int64 * 4)()
    _check(lib().pth_num_entities(i, j, k, nrefine, out))
    return tuple(out)


def num_pdofs(i, j, k, nrefine, order):
    out = C.c_int64()
    _check(lib().pth_num_pdofs(i, j, k, nrefine, order, C.byref(out)))
    return out.value


def cube_sizing(target_dofs, total, dofs_per_node, order, num_processes=1):
    """(Nx, Ny, Nz, r) exactly as create_cube_mesh picks them (src/mesh.cpp:78-151)."""
    out = (C.c_int64 * 4)()
    _check(lib().pth_cube_sizing(target_dofs, int(bool(total)), dofs_per_node, order,
                                 num_processes, out))
    return tuple(out)


class Problem:
    """Mesh slab + function space + BC + RHS + sparsity pattern for one rank (host memory)."""

    def __init__(self, problem_type: str, order: int, nx: int, ny: int, nz: int,
                 rank: int = 0, nranks: int = 1, with_dofmap: bool = True, renumber: str | None = None,
                 seed: int = 0):
        """with_dofmap=False: sizes, halo lists and exterior facets only (the arrays are generated
        on the device by Context.set_problem_on_device). renumber = "rcm" | "random": owned dofs
        renumbered like a DOLFINx dofmap is (single rank; include/ptb200_host.h)."""
        self._h = C.c_void_p()
        create = lib().pth_problem_create if with_dofmap else lib().pth_problem_create_sizes_only
        _check(create(problem_type.encode(), order, nx, ny, nz, rank, nranks, C.byref(self._h)))
        if renumber:
            _check(lib().pth_problem_renumber(self._h, renumber.encode(), seed))
        self.renumber = renumber
        self.problem_type = problem_type
        for name in SCALARS:
            v = C.c_int64()
            _check(lib().pth_problem_scalar(self._h, name.encode(), C.byref(v)))
            setattr(self, name, v.value)
        self._views = {}

    def __getitem__(self, name: str) -> np.ndarray:
        """Zero-copy numpy view of a named array (valid while this object lives)."""
        if name not in self._views:
            p, n, dt = C.c_void_p(), C.c_int64(), C.c_int()
            _check(lib().pth_problem_array(self._h, name.encode(), C.byref(p), C.byref(n),
                                           C.byref(dt)))
            dtype = np.dtype(_DT[dt.value])
            if n.value == 0:
                a = np.zeros(0, dtype=dtype)
            else:
                buf = (C.c_char * (n.value * dtype.itemsize)).from_address(p.value)
                a = np.frombuffer(buf, dtype=dtype)
                a.flags.writeable = False
            self._views[name] = a
        return self._views[name]

    def close(self):
        if self._h:
            self._views.clear()
            lib().pth_problem_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
