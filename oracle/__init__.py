"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the reference's hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this package. It is the checker, never the product: the product path is
``performance-test_b200`` (CUDA, no CPU fallback).

Pinning: the reference has no golden vectors or KATs and its assembly lives
in DOLFINx/FFCx/PETSc, which cannot be built here -- but its OWN code on this path can: src/cg.h
(unchanged), the sizing arithmetic of src/mesh.cpp, the BC predicates / source lambdas and
pack_fn/unpack_fn are compiled from /root/reference into oracle/_ref/libref.so (oracle/ref/,
loader oracle/ref.py) and this restatement is checked against them bit for bit
(tests/test_ref_pin.py, tests/golden/ref_cg.json, tests/golden/sizing.json). The element kernels,
assembly loop and the Jacobi extension of the CG loop remain "parity unpinned": restated from
published DOLFINx/FFCx semantics and pinned by analytic KATs only (tests/test_oracle_kats.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import tables as _tables

_HERE = os.path.dirname(os.path.abspath(__file__))
_libs = {}
_tab_cache = {}


def build():
    r = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    from . import ref as _ref
    _ref.build()   # oracle/_ref from /root/reference when it is present (no-op on the GPU box)


def _lib(fast=False):
    key = "fast" if fast else "strict"
    if key not in _libs:
        path = os.path.join(_HERE, "_build", "liboracle_fast.so" if fast else "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_max_threads.restype = C.c_int
        L.orc_cg.restype = C.c_int
        _libs[key] = L
    return _libs[key]


def max_threads():
    return _lib().orc_max_threads()


def element_tables(order):
    if order not in _tab_cache:
        _tab_cache[order] = _tables.element_tables(order)
    return _tab_cache[order]


def _p(a, dtype):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=dtype)
    return a, a.ctypes.data_as(C.c_void_p)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _arr(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


PROBLEMS = {"poisson": 0, "cgpoisson": 0, "elasticity": 1}


def bc_marker(P):
    m = np.zeros(P.n_owned + P.n_ghost, dtype=np.int8)
    m[P["bc_dofs"]] = 1
    return m


def assemble_matrix(P, slot=None, nthreads=1, fast=False, set_diagonal=True):
    """ZZZ Assemble matrix (src/poisson_problem.cpp:125-139). Returns block-CSR values
    [nnz*bs*bs] on the pattern P['rowptr'], P['cols'] (owned rows)."""
    T = element_tables(P.order)
    bs = P.bs
    x = _arr(P["x"], np.float64)
    xd = _arr(P["x_dofmap"], np.int32)
    dm = _arr(P["dofmap"], np.int32)
    rp = _arr(P["rowptr"], np.int64)
    cl = _arr(P["cols"], np.int32)
    bcd = _arr(P["bc_dofs"], np.int32)
    mk = bc_marker(P)
    vals = np.zeros(len(cl) * bs * bs, dtype=np.float64)
    sl = None if slot is None else _arr(slot, np.int64)
    L = _lib(fast)
    rc = L.orc_assemble_matrix(
        C.c_int(PROBLEMS[P.problem_type]), C.c_int(T["nd"]), C.c_int(bs), C.c_int(T["nq_a"]),
        _ptr(T["w_a"]), _ptr(T["dphi_a"]), _ptr(T["dgeo"]), C.c_int64(P.n_cells), _ptr(x),
        _ptr(xd), _ptr(dm), C.c_int32(P.n_owned), _ptr(rp), _ptr(cl), _ptr(mk),
        None if sl is None else _ptr(sl), _ptr(vals), C.c_int(nthreads))
    if rc != 0:
        raise RuntimeError("oracle: (row, col) missing from the sparsity pattern")
    if set_diagonal:
        rc = L.orc_set_diagonal(C.c_int(bs), C.c_int32(P.n_owned), _ptr(rp), _ptr(cl),
                                C.c_int32(len(bcd)), _ptr(bcd), _ptr(vals))
        if rc != 0:
            raise RuntimeError("oracle: diagonal missing from the sparsity pattern")
    return vals


def assemble_vector(P, fast=False):
    """ZZZ Assemble vector (src/poisson_problem.cpp:146-157), owned rows. apply_lifting is
    omitted: g = u0 = 0 makes it a numerical no-op."""
    T = element_tables(P.order)
    bs = P.bs
    x = _arr(P["x"], np.float64)
    xd = _arr(P["x_dofmap"], np.int32)
    dm = _arr(P["dofmap"], np.int32)
    f = _arr(P["f"], np.float64)
    g = _arr(P["g"], np.float64) if len(P["g"]) else None
    fc = _arr(P["facet_cells"], np.int32)
    fl = _arr(P["facet_local"], np.int32)
    bcd = _arr(P["bc_dofs"], np.int32)
    b = np.zeros(P.n_owned * bs, dtype=np.float64)
    _lib(fast).orc_assemble_vector(
        C.c_int(T["nd"]), C.c_int(bs), C.c_int(T["nq_l"]), _ptr(T["w_l"]), _ptr(T["phi_l"]),
        C.c_int(T["nq_f"]), _ptr(T["w_f"]), _ptr(T["phi_f"]), _ptr(T["dgeo"]),
        _ptr(T["facet_t"]), C.c_int64(P.n_cells), _ptr(x), _ptr(xd), _ptr(dm),
        C.c_int32(P.n_owned), _ptr(f), None if g is None else _ptr(g), C.c_int64(len(fc)),
        _ptr(fc), _ptr(fl), C.c_int32(len(bcd)), _ptr(bcd), _ptr(b))
    return b


def spmv(bs, n_rows, rowptr, cols, vals, p, nthreads=1, fast=False):
    y = np.zeros(n_rows * bs, dtype=np.float64)
    p = _arr(p, np.float64)
    _lib(fast).orc_spmv(C.c_int(bs), C.c_int32(n_rows), _ptr(_arr(rowptr, np.int64)),
                        _ptr(_arr(cols, np.int32)), _ptr(_arr(vals, np.float64)), _ptr(p),
                        _ptr(y), C.c_int(nthreads))
    return y


def cg(bs, n_rows, rowptr, cols, vals, b, x0=None, kmax=50, rtol=1e-8, precond="none",
       nthreads=1, fast=False):
    """linalg::cg (src/cg.h:38-86) [+ Jacobi]; single partition. Returns (x, iterations, rel_res)."""
    x = np.zeros(n_rows * bs) if x0 is None else np.array(x0, dtype=np.float64)
    rel = C.c_double()
    rp, cl, vl, bb = (_arr(rowptr, np.int64), _arr(cols, np.int32), _arr(vals, np.float64),
                      _arr(b, np.float64))
    k = _lib(fast).orc_cg(C.c_int(bs), C.c_int32(n_rows), _ptr(rp), _ptr(cl), _ptr(vl), _ptr(bb),
                          _ptr(x), C.c_int(kmax), C.c_double(rtol),
                          C.c_int({"none": 0, "jacobi": 1}[precond]), C.byref(rel),
                          C.c_int(nthreads))
    return x, k, rel.value


class _Part(C.Structure):
    _fields_ = [("bs", C.c_int32), ("n_owned", C.c_int32), ("n_ghost", C.c_int32),
                ("n_nbr", C.c_int32), ("rowptr", C.c_void_p), ("cols", C.c_void_p),
                ("vals", C.c_void_p), ("b", C.c_void_p), ("x", C.c_void_p),
                ("nbr_ranks", C.c_void_p), ("send_displ", C.c_void_p),
                ("local_indices", C.c_void_p), ("recv_displ", C.c_void_p),
                ("remote_indices", C.c_void_p)]


def cg_partitioned(problems, mats, rhs, kmax=50, rtol=1e-8, precond="none", fast=False):
    """The solve on len(problems) partitions, one thread per partition (the analogue of
    `mpirun -np P`): halo update of p with the Scatterer lists, dot products reduced in rank
    order. problems: host.Problem per rank; mats / rhs: their owned-row matrices and vectors.
    Returns ([x per rank, owned + ghost], iterations, rel_res)."""
    keep, arr, xs = [], (_Part * len(problems))(), []
    for q, (P, A, b) in enumerate(zip(problems, mats, rhs)):
        x = np.zeros((P.n_owned + P.n_ghost) * P.bs)
        fields = [_arr(P["rowptr"], np.int64), _arr(P["cols"], np.int32), _arr(A, np.float64),
                  _arr(b, np.float64), x]
        halo = [_arr(P[k] if len(P[k]) else np.zeros(1 if "displ" in k else 0), np.int32)
                for k in ("nbr_ranks", "send_displ", "local_indices", "recv_displ", "remote_indices")]
        keep += fields + halo
        xs.append(x)
        arr[q] = _Part(P.bs, P.n_owned, P.n_ghost, P.n_nbr, *[_ptr(f) for f in fields],
                       *[_ptr(h) for h in halo])
    k, rel = C.c_int(), C.c_double()
    rc = _lib(fast).orc_cg_partitioned(C.c_int(len(problems)), arr, C.c_int(kmax), C.c_double(rtol),
                                       C.c_int({"none": 0, "jacobi": 1}[precond]), C.byref(k),
                                       C.byref(rel))
    if rc != 0:
        raise RuntimeError("oracle: inconsistent halo lists between partitions")
    return xs, k.value, rel.value


def assemble_partitions(problems, fast=False):
    """ZZZ Assemble matrix / vector of every partition at once, one thread per partition (each runs
    the sequential cell loop of its rank, like the ranks of an MPI job). Returns (mats, rhs,
    seconds_matrix, seconds_vector)."""
    import time
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=len(problems)) as ex:
        t0 = time.perf_counter()
        mats = list(ex.map(lambda P: assemble_matrix(P, nthreads=1, fast=fast), problems))
        t1 = time.perf_counter()
        rhs = list(ex.map(lambda P: assemble_vector(P, fast=fast), problems))
        t2 = time.perf_counter()
    return mats, rhs, t1 - t0, t2 - t1
