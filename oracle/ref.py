"""TEST INFRASTRUCTURE -- ctypes loader of oracle/_ref/libref.so: the reference's own code
(src/cg.h unchanged; the sizing arithmetic of src/mesh.cpp; the BC predicates and source-term
lambdas of src/poisson_problem.cpp / src/elasticity_problem.cpp; pack_fn / unpack_fn of
src/cgpoisson_problem.cpp) compiled by oracle/ref/Makefile from /root/reference where it lies.

/root/reference does not exist on the GPU box; the built .so travels there (oracle/_ref/ is
git-ignored, not gpurun-ignored). Only tests/, smoke() and bench.py's reference leg may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_ROOT = os.environ.get("PTB_REFERENCE_ROOT", "/root/reference")
_libs = {}


def _path(fast=False):
    return os.path.join(_HERE, "_ref", "libref_fast.so" if fast else "libref.so")


def build():
    """Compile oracle/_ref from the reference's sources (no-op without /root/reference: the GPU
    box only uses the prebuilt files)."""
    if not os.path.isdir(os.path.join(_REF_ROOT, "src")):
        return False
    r = subprocess.run(["make", "-C", os.path.join(_HERE, "ref"), f"REF={_REF_ROOT}", "all"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle/_ref build failed:\n" + r.stdout + r.stderr)
    return True


def available():
    if not os.path.exists(_path()):
        try:
            build()
        except RuntimeError:
            return False
    return os.path.exists(_path())


class Part(C.Structure):
    _fields_ = [("bs", C.c_int32), ("n_owned", C.c_int32), ("n_ghost", C.c_int32),
                ("n_nbr", C.c_int32), ("rowptr", C.c_void_p), ("cols", C.c_void_p),
                ("vals", C.c_void_p), ("b", C.c_void_p), ("x", C.c_void_p),
                ("nbr_ranks", C.c_void_p), ("send_displ", C.c_void_p),
                ("local_indices", C.c_void_p), ("recv_displ", C.c_void_p),
                ("remote_indices", C.c_void_p)]


def lib(fast=False):
    key = bool(fast)
    if key not in _libs:
        if not available():
            raise RuntimeError("oracle/_ref/libref.so is missing and /root/reference is not here "
                               "to build it from")
        L = C.CDLL(_path(fast))
        L.ref_num_entities.argtypes = [C.c_int64] * 3 + [C.c_int, C.POINTER(C.c_int64)]
        L.ref_num_entities.restype = None
        L.ref_num_pdofs.argtypes = [C.c_int64] * 3 + [C.c_int, C.c_int, C.POINTER(C.c_int64)]
        L.ref_cube_sizing.argtypes = [C.c_uint64, C.c_int, C.c_uint64, C.c_int, C.c_int,
                                      C.POINTER(C.c_int64)]
        L.ref_bc_marker.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
        L.ref_bc_marker.restype = None
        L.ref_poisson_source.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_poisson_source.restype = None
        L.ref_elasticity_source.argtypes = [C.c_int64, C.c_void_p, C.c_void_p]
        L.ref_elasticity_source.restype = None
        L.ref_pack.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.ref_pack.restype = None
        L.ref_unpack.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
        L.ref_unpack.restype = None
        L.ref_axpy.argtypes = [C.c_int32, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_axpy.restype = None
        L.ref_cg.argtypes = [C.c_int, C.POINTER(Part), C.c_int, C.c_double, C.POINTER(C.c_int)]
        _libs[key] = L
    return _libs[key]


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


# ---- src/mesh.cpp ------------------------------------------------------------------------------
def num_entities(i, j, k, nrefine=0):
    out = (C.c_int64 * 4)()
    lib().ref_num_entities(i, j, k, nrefine, out)
    return tuple(out)


def num_pdofs(i, j, k, nrefine, order):
    out = C.c_int64()
    if lib().ref_num_pdofs(i, j, k, nrefine, order, C.byref(out)) != 0:
        raise RuntimeError("Order not supported")
    return out.value


def cube_sizing(target_dofs, total, dofs_per_node, order, num_processes=1):
    out = (C.c_int64 * 4)()
    if lib().ref_cube_sizing(target_dofs, int(bool(total)), dofs_per_node, order, num_processes,
                             out) != 0:
        raise RuntimeError("Order not supported")
    return tuple(out)


# ---- problem data: points [n][3] in, the lambdas see them as x(i, p) ----------------------------
def _x3n(points):
    pts = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    return np.ascontiguousarray(pts.T), len(pts)


def bc_marker(problem_type, points):
    x, n = _x3n(points)
    m = np.zeros(n, dtype=np.int8)
    lib().ref_bc_marker(int(problem_type == "elasticity"), n, _ptr(x), _ptr(m))
    return m.astype(bool)


def poisson_source(points, fast=False):
    x, n = _x3n(points)
    f, g = np.zeros(n), np.zeros(n)
    lib(fast).ref_poisson_source(n, _ptr(x), _ptr(f), _ptr(g))
    return f, g


def elasticity_source(points, fast=False):
    """[n][3] (blocked, the layout of f->x()->array())."""
    x, n = _x3n(points)
    f = np.zeros((3, n))
    lib(fast).ref_elasticity_source(n, _ptr(x), _ptr(f))
    return np.ascontiguousarray(f.T)


# ---- src/cgpoisson_problem.cpp:32-44 -------------------------------------------------------------
def pack(values, idx):
    values = np.ascontiguousarray(values, dtype=np.float64)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    out = np.zeros(len(idx))
    lib().ref_pack(len(idx), _ptr(idx), _ptr(values), len(values), _ptr(out))
    return out


def unpack(buf, idx, out, op="plus"):
    buf = np.ascontiguousarray(buf, dtype=np.float64)
    idx = np.ascontiguousarray(idx, dtype=np.int32)
    out = np.array(out, dtype=np.float64)
    lib().ref_unpack(len(idx), _ptr(idx), _ptr(buf), len(out), _ptr(out),
                     {"plus": 0, "overwrite": 1}[op])
    return out


# ---- src/cg.h -------------------------------------------------------------------------------------
def axpy(alpha, x, y):
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    r = np.zeros_like(x)
    lib().ref_axpy(len(x), alpha, _ptr(x), _ptr(y), _ptr(r))
    return r


def cg(parts, kmax=50, rtol=1e-8, fast=False):
    """linalg::cg (src/cg.h:38-86, compiled unchanged) over len(parts) partitions (threads).

    parts: dicts with bs, n_owned, n_ghost, rowptr, cols, vals, b [(owned+ghost)*bs], optional x0,
    and for more than one partition nbr_ranks, send_displ, local_indices, recv_displ,
    remote_indices (the lists of ptb_set_halo). Returns ([x per part], iterations)."""
    keep, arr = [], (Part * len(parts))()
    xs = []
    for q, d in enumerate(parts):
        bs, no, ng = int(d["bs"]), int(d["n_owned"]), int(d["n_ghost"])
        nl = (no + ng) * bs
        x = np.zeros(nl) if d.get("x0") is None else np.array(d["x0"], dtype=np.float64)
        b = np.ascontiguousarray(d["b"], dtype=np.float64)
        assert len(x) == nl and len(b) == nl
        rp = np.ascontiguousarray(d["rowptr"], dtype=np.int64)
        cl = np.ascontiguousarray(d["cols"], dtype=np.int32)
        vl = np.ascontiguousarray(d["vals"], dtype=np.float64)
        halo = [np.ascontiguousarray(d.get(k, np.zeros(1 if "displ" in k else 0)), dtype=np.int32)
                for k in ("nbr_ranks", "send_displ", "local_indices", "recv_displ",
                          "remote_indices")]
        keep += [x, b, rp, cl, vl] + halo
        xs.append(x)
        arr[q] = Part(bs, no, ng, len(halo[0]), _ptr(rp), _ptr(cl), _ptr(vl), _ptr(b), _ptr(x),
                      *[_ptr(h) for h in halo])
    iters = (C.c_int * len(parts))()
    if lib(fast).ref_cg(len(parts), arr, kmax, rtol, iters) != 0:
        raise RuntimeError("ref_cg: inconsistent halo lists")
    its = list(iters)
    assert all(i == its[0] for i in its), its
    return xs, its[0]
