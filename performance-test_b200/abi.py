"""ctypes binding of libptb200.so (include/ptb200.h): the drop-in C-ABI over the sm_100a kernels.

No CPU fallback: ``Context()`` raises when the library or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

_lib = None

POISSON, ELASTICITY = 0, 1
PC_NONE, PC_JACOBI = 0, 1
STAGE_ASSEMBLE_MATRIX, STAGE_ASSEMBLE_VECTOR, STAGE_SOLVE, STAGE_SPMV = 0, 1, 2, 3
KERNEL_SPMV, KERNEL_CG_UPDATE, KERNEL_CG_DIRECTION, KERNEL_ASSEMBLE_MATRIX, \
    KERNEL_ASSEMBLE_VECTOR = range(5)
PROBLEMS = {"poisson": POISSON, "cgpoisson": POISSON, "elasticity": ELASTICITY}
PRECOND = {"none": PC_NONE, "jacobi": PC_JACOBI}


def declared_symbols():
    """Every function include/ptb200.h declares (used by the CPU-side export test)."""
    from . import INCLUDE_DIR
    src = open(os.path.join(INCLUDE_DIR, "ptb200.h")).read()
    src += open(os.path.join(INCLUDE_DIR, "ptb200_debug.h")).read()   # test hooks, same library
    return sorted(set(re.findall(r"\b(ptb_[a-z0-9_]+)\s*\(", src)))


def lib():
    global _lib
    if _lib is None:
        from . import ABI_LIB
        if not os.path.exists(ABI_LIB):
            raise RuntimeError(f"{ABI_LIB} is missing: run __graft_entry__.build() / make "
                               "(there is no CPU fallback)")
        L = C.CDLL(ABI_LIB)
        vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
        L.ptb_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.ptb_destroy.argtypes = [vp]
        L.ptb_destroy.restype = None
        L.ptb_last_error.argtypes = [vp]
        L.ptb_last_error.restype = C.c_char_p
        L.ptb_set_stream.argtypes = [vp, vp]
        L.ptb_set_mesh.argtypes = [vp, i64, vp, i64, vp]
        L.ptb_update_geometry.argtypes = [vp, vp]
        L.ptb_set_space.argtypes = [vp, C.c_int, C.c_int, C.c_int, i32, i32, vp]
        L.ptb_set_pattern.argtypes = [vp, vp, vp]
        L.ptb_create_box.argtypes = [vp, C.c_int, C.c_int, C.c_int, i64, i64, i64, C.c_int, C.c_int, vp]
        L.ptb_get_dof_coordinates.argtypes = [vp, vp]
        L.ptb_get_mesh.argtypes = [vp, vp, vp]
        L.ptb_get_dofmap.argtypes = [vp, vp]
        L.ptb_build_pattern.argtypes = [vp, C.POINTER(i64)]
        L.ptb_get_pattern.argtypes = [vp, vp, vp]
        L.ptb_locate_bc.argtypes = [vp, C.POINTER(i32)]
        L.ptb_get_bc.argtypes = [vp, vp]
        L.ptb_interpolate_source.argtypes = [vp, vp]
        L.ptb_get_source.argtypes = [vp, vp, vp]
        L.ptb_set_bc.argtypes = [vp, i32, vp]
        L.ptb_set_exterior_facets.argtypes = [vp, i64, vp, vp]
        L.ptb_set_source.argtypes = [vp, vp, vp]
        L.ptb_set_halo.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp]
        L.ptb_nccl_unique_id.argtypes = [vp]
        L.ptb_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
        L.ptb_peer_export.argtypes = [vp, vp]
        L.ptb_peer_connect.argtypes = [vp, C.c_int, C.c_int, vp, vp]
        L.ptb_assemble_matrix.argtypes = [vp]
        L.ptb_assemble_vector.argtypes = [vp]
        L.ptb_cg_solve.argtypes = [vp, C.c_int, dbl, C.c_int, C.POINTER(C.c_int), C.POINTER(dbl)]
        L.ptb_set_operator_mode.argtypes = [vp, C.c_int]
        L.ptb_set_cg_persistent.argtypes = [vp, C.c_int]
        L.ptb_apply_operator.argtypes = [vp, vp, vp]
        L.ptb_set_rhs.argtypes = [vp, vp]
        L.ptb_set_initial_guess.argtypes = [vp, vp]
        L.ptb_get_matrix_values.argtypes = [vp, vp]
        L.ptb_get_diagonal_inverse.argtypes = [vp, vp]
        L.ptb_get_rhs.argtypes = [vp, vp]
        L.ptb_get_solution.argtypes = [vp, vp]
        L.ptb_solution_norm.argtypes = [vp, C.POINTER(dbl)]
        L.ptb_build_cell_slot_map.argtypes = [i64, C.c_int, vp, i32, vp, vp, vp]
        L.ptb_debug_layout_roundtrip.argtypes = [i32, i64, vp, vp, vp, C.POINTER(dbl)]
        L.ptb_get_slot_offsets.argtypes = [vp, C.POINTER(i64), vp, vp, vp]
        L.ptb_get_p1_maps.argtypes = [vp, vp, vp, vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ptb_get_p1_rings.argtypes = [vp, vp, vp, vp, C.POINTER(C.c_int)]
        L.ptb_debug_star_walk.argtypes = [i64, vp, i32, vp, vp, vp, C.POINTER(dbl)]
        L.ptb_debug_star_walk_single.argtypes = [i64, vp, i32, vp, vp, vp, vp]
        L.ptb_debug_facet_rows.argtypes = [i64, vp, vp, vp, C.c_int, C.c_int, i32, C.POINTER(i32),
                                           C.POINTER(i32), vp, vp, vp]
        L.ptb_debug_facet_rows_gathered.argtypes = L.ptb_debug_facet_rows.argtypes
        L.ptb_debug_pk_layout.argtypes = [i64, C.c_int, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.ptb_debug_slice_order.argtypes = [i32, vp, vp, vp, C.POINTER(i32)]
        L.ptb_debug_compressed_columns.argtypes = [i32, i64, vp, vp, vp, vp, vp]
        L.ptb_debug_p1_rings.argtypes = [i64, vp, i32, vp, vp, vp, vp, vp]
        L.ptb_debug_p1_layout.argtypes = [i64, vp, i32, vp, vp, C.POINTER(C.c_int), vp, vp, vp, vp, vp, vp, vp]
        L.ptb_time_kernel.argtypes = [vp, C.c_int, C.c_int, C.POINTER(dbl)]
        L.ptb_stage_ms.argtypes = [vp, C.c_int]
        L.ptb_stage_ms.restype = dbl
        L.ptb_launch_count.argtypes = [vp]
        L.ptb_launch_count.restype = i64
        L.ptb_cols_explicit_fraction.argtypes = [vp]
        L.ptb_cols_explicit_fraction.restype = dbl
        L.ptb_spmv_stored_entries.argtypes = [vp]
        L.ptb_spmv_stored_entries.restype = i64
        L.ptb_device_bytes.argtypes = [vp]
        L.ptb_device_bytes.restype = i64
        _lib = L
    return _lib


def _a(x, dtype):
    return np.ascontiguousarray(x, dtype=dtype)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def build_cell_slot_map(dofmap, nd, n_owned, rowptr, cols):
    """Host-only (no GPU): the uncompressed cell -> CSR-slot map."""
    dm, rp, cl = _a(dofmap, np.int32), _a(rowptr, np.int64), _a(cols, np.int32)
    n_cells = len(dm) // nd
    out = np.empty(n_cells * nd * nd, dtype=np.int64)
    rc = lib().ptb_build_cell_slot_map(n_cells, nd, _ptr(dm), n_owned, _ptr(rp), _ptr(cl),
                                       _ptr(out))
    if rc != 0:
        raise RuntimeError(lib().ptb_last_error(None).decode())
    return out


def star_walk(dofmap, n_owned, rowptr, cols):
    """Host-only: P1 star-walk step words per (row, step), row-major; returns (words, loads/step)."""
    dm, rp, cl = _a(dofmap, np.int32), _a(rowptr, np.int64), _a(cols, np.int32)
    n_cells = len(dm) // 4
    n_pairs = int(np.count_nonzero(dm < n_owned))
    out = np.zeros(n_pairs, dtype=np.uint32)
    lps = C.c_double()
    rc = lib().ptb_debug_star_walk(n_cells, _ptr(dm), n_owned, _ptr(rp), _ptr(cl), _ptr(out),
                                   C.byref(lps))
    if rc != 0:
        raise RuntimeError(lib().ptb_last_error(None).decode())
    return out, lps.value


def star_walk_single(dofmap, n_owned, rowptr, cols):
    """Host-only: the one-vertex-per-step walk; returns (step_ptr [n_owned+1], words)."""
    dm, rp, cl = _a(dofmap, np.int32), _a(rowptr, np.int64), _a(cols, np.int32)
    n_cells = len(dm) // 4
    ptr = np.zeros(n_owned + 1, dtype=np.int64)
    for words in (None, "alloc"):
        if words is not None:
            words = np.zeros(int(ptr[-1]), dtype=np.uint32)
        rc = lib().ptb_debug_star_walk_single(n_cells, _ptr(dm), n_owned, _ptr(rp), _ptr(cl),
                                              _ptr(ptr), _ptr(words))
        if rc != 0:
            raise RuntimeError(lib().ptb_last_error(None).decode())
    return ptr, words


def p1_layout(dofmap, n_owned, rowptr, cols):
    """Host-only: the SELL-32 arrays of the P1 walk kernels as a dict (device order)."""
    dm, rp, cl = _a(dofmap, np.int32), _a(rowptr, np.int64), _a(cols, np.int32)
    n_cells, ns = len(dm) // 4, (n_owned + 31) // 32
    mw = C.c_int()
    off = {k: np.zeros(ns + 1, dtype=np.int64) for k in ("mat_off", "adj_off", "walk1_off")}
    data = {}
    for fill in (False, True):
        if fill:
            data = {"cols": np.zeros(int(off["mat_off"][-1]), dtype=np.int32),
                    "walk": np.zeros(int(off["adj_off"][-1]), dtype=np.uint32),
                    "walk1": np.zeros(int(off["walk1_off"][-1]), dtype=np.uint32),
                    "adjrot": np.zeros(int(off["adj_off"][-1]), dtype=np.uint32)}
        rc = lib().ptb_debug_p1_layout(n_cells, _ptr(dm), n_owned, _ptr(rp), _ptr(cl), C.byref(mw),
                                       _ptr(off["mat_off"]), _ptr(off["adj_off"]),
                                       _ptr(off["walk1_off"]), _ptr(data.get("cols")),
                                       _ptr(data.get("walk")), _ptr(data.get("walk1")),
                                       _ptr(data.get("adjrot")))
        if rc != 0:
            raise RuntimeError(lib().ptb_last_error(None).decode())
    return dict(off, **data, max_w=mw.value, n_slices=ns)


def p1_rings(dofmap, n_owned, rowptr, cols, n_sell_entries):
    """Host-only: (ring_off, ring_ns, ring) of the column-major elasticity kernel (layout.h build_rings)
    for a pattern whose SELL-32 layout has n_sell_entries padded entries (mat_off[-1])."""
    dm, rp, cl = _a(dofmap, np.int32), _a(rowptr, np.int64), _a(cols, np.int32)
    ring_off = np.zeros((n_owned + 31) // 32 + 1, dtype=np.int64)
    ring_ns = np.zeros(n_sell_entries // 32, dtype=np.uint8)
    ring = None
    for fill in (False, True):
        if fill:
            ring = np.zeros(max(int(ring_off[-1]), 1), dtype=np.uint32)
        rc = lib().ptb_debug_p1_rings(len(dm) // 4, _ptr(dm), n_owned, _ptr(rp), _ptr(cl), _ptr(ring_off),
                                      _ptr(ring_ns), _ptr(ring))
        if rc != 0:
            raise RuntimeError(lib().ptb_last_error(None).decode())
    return ring_off, ring_ns, ring


def compressed_columns(n_rows, n_cols, rowptr, cols, n_sell_entries):
    """Host-only: (cdelta, xoff, colsx) of the scalar SpMV for a pattern whose SELL-32 layout has
    n_sell_entries padded entries (mat_off[-1])."""
    rp, cl = _a(rowptr, np.int64), _a(cols, np.int32)
    cdelta = np.zeros(n_sell_entries // 32, dtype=np.int32)
    xoff = np.zeros((n_rows + 31) // 32 + 1, dtype=np.int64)
    colsx = None
    for fill in (False, True):
        if fill:
            colsx = np.zeros(max(int(xoff[-1]), 1), dtype=np.int32)
        rc = lib().ptb_debug_compressed_columns(n_rows, n_cols, _ptr(rp), _ptr(cl), _ptr(cdelta),
                                                _ptr(xoff), _ptr(colsx))
        if rc != 0:
            raise RuntimeError(lib().ptb_last_error(None).decode())
    return cdelta, xoff, colsx


def slice_order(n_rows, rowptr, cols):
    """Host-only: (order, n_interior) of the operator kernels' slice visiting order."""
    rp, cl = _a(rowptr, np.int64), _a(cols, np.int32)
    order = np.zeros((n_rows + 31) // 32, dtype=np.int32)
    ni = C.c_int32()
    if lib().ptb_debug_slice_order(n_rows, _ptr(rp), _ptr(cl), _ptr(order), C.byref(ni)) != 0:
        raise RuntimeError(lib().ptb_last_error(None).decode())
    return order, ni.value


def balance_plan(mat_off, order, n_interior, grid, npull=-1):
    """Host-only: (ounit, begin) of the balanced operator split (layout.h build_balance_plan)."""
    mo, od = _a(mat_off, np.int64), _a(order, np.int32)
    S = len(od)
    ounit = np.zeros(S + 1, dtype=np.int32)
    begin = np.zeros(grid + 2, dtype=np.int32)
    nb = C.c_int32()
    if lib().ptb_debug_balance_plan(S, _ptr(mo), _ptr(od), n_interior, grid, npull, _ptr(ounit), _ptr(begin),
                                    C.byref(nb)) != 0:
        raise RuntimeError(lib().ptb_last_error(None).decode())
    return ounit, begin[: nb.value].copy()


def pk_layout(dofmap, nd, n_owned, rowptr, cols):
    """Host-only: SELL-32 arrays of the P2/P3 assembly kernels + row-length bins, as a dict."""
    dm, rp, cl = _a(dofmap, np.int32), _a(rowptr, np.int64), _a(cols, np.int32)
    n_cells, ns = len(dm) // nd, (n_owned + 31) // 32
    info = np.zeros(4, dtype=np.int32)
    d = {"mat_off": np.zeros(ns + 1, np.int64), "adj_off": np.zeros(ns + 1, np.int64),
         "bin_off": np.zeros(17, np.int32), "bin_w": np.zeros(16, np.int32)}
    data = {}
    for fill in (False, True):
        if fill:
            data = {"cols": np.zeros(int(d["mat_off"][-1]), np.int32),
                    "adj": np.zeros(int(d["adj_off"][-1]), np.uint32),
                    "adjso": np.zeros(int(d["adj_off"][-1]) * int(info[2]), np.uint32),
                    "bin_slices": np.zeros(ns, np.int32)}
        rc = lib().ptb_debug_pk_layout(n_cells, nd, _ptr(dm), n_owned, _ptr(rp), _ptr(cl), _ptr(info),
                                       _ptr(d["mat_off"]), _ptr(d["adj_off"]), _ptr(d["bin_off"]),
                                       _ptr(d["bin_w"]), _ptr(data.get("cols")), _ptr(data.get("adj")),
                                       _ptr(data.get("adjso")), _ptr(data.get("bin_slices")))
        if rc != 0:
            raise RuntimeError(lib().ptb_last_error(None).decode())
    return dict(d, **data, max_w=int(info[0]), so_bits=int(info[1]), so_words=int(info[2]),
                n_bins=int(info[3]), n_slices=ns)


def facet_rows(facet_cells, facet_local, dofmap, nd, order, n_rows, gathered=False):
    """Host-only: (row_ids, row_ptr, ent) of the boundary-facet gather (ent = int32 pairs). With
    gathered, `dofmap` holds only the rows of the facets' cells ([n_facets * nd])."""
    fc, fl, dm = _a(facet_cells, np.int32), _a(facet_local, np.int32), _a(dofmap, np.int32)
    ids, ptr = np.zeros(max(n_rows, 1), np.int32), np.zeros(n_rows + 1, np.int32)
    ent = np.zeros(max(20 * len(fc), 2), np.int32)
    nf, ne = C.c_int32(), C.c_int32()
    fn = lib().ptb_debug_facet_rows_gathered if gathered else lib().ptb_debug_facet_rows
    rc = fn(len(fc), _ptr(fc), _ptr(fl), _ptr(dm), nd, order, n_rows,
            C.byref(nf), C.byref(ne), _ptr(ids), _ptr(ptr), _ptr(ent))
    if rc != 0:
        raise RuntimeError(lib().ptb_last_error(None).decode())
    return ids[:nf.value].copy(), ptr[:nf.value + 1].copy(), ent[:2 * ne.value].copy()


def layout_roundtrip(n_rows, n_cols, rowptr, cols):
    """Host-only: SELL-32 + column compression encode/decode; returns (cols decoded, explicit fraction)."""
    rp, cl = _a(rowptr, np.int64), _a(cols, np.int32)
    out = np.full(len(cl), -1, dtype=np.int32)
    frac = C.c_double()
    rc = lib().ptb_debug_layout_roundtrip(n_rows, n_cols, _ptr(rp), _ptr(cl), _ptr(out),
                                          C.byref(frac))
    if rc != 0:
        raise RuntimeError(lib().ptb_last_error(None).decode())
    return out, frac.value


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    if lib().ptb_nccl_unique_id(buf) != 0:
        raise RuntimeError(lib().ptb_last_error(None).decode())
    return buf.raw


class Context:
    """One GPU / one rank. Mirrors the C-ABI one to one."""

    def __init__(self, device: int = 0, stream=None):
        self._h = C.c_void_p()
        if lib().ptb_create(device, C.byref(self._h)) != 0:
            raise RuntimeError(lib().ptb_last_error(None).decode())
        if stream is not None:
            self._check(lib().ptb_set_stream(self._h, C.c_void_p(stream)))
        self.P = None

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(lib().ptb_last_error(self._h).decode())

    def close(self):
        if self._h:
            lib().ptb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- setup ---------------------------------------------------------------------------
    def set_problem(self, P, source=True, build_pattern=False, device_data=False):
        """Upload everything a host problem (host.Problem or the oracle's RefProblem) holds. With
        build_pattern the sparsity pattern is built on the device from the dofmap
        (ptb_build_pattern) instead of being uploaded; with device_data the Dirichlet dofs and the
        source terms are evaluated on the device (ptb_locate_bc, ptb_interpolate_source)."""
        self.P = P
        self.bs, self.nd = P.bs, P.nd
        self.n_owned, self.n_ghost, self.nnz = P.n_owned, P.n_ghost, P.nnz
        x, xd = _a(P["x"], np.float64), _a(P["x_dofmap"], np.int32)
        self.n_vertices, self.n_cells = len(x) // 3, len(xd) // 4
        self._check(lib().ptb_set_mesh(self._h, len(x) // 3, _ptr(x), len(xd) // 4, _ptr(xd)))
        dm = _a(P["dofmap"], np.int32)
        self._check(lib().ptb_set_space(self._h, PROBLEMS[P.problem_type], P.order, P.bs,
                                        P.n_owned, P.n_ghost, _ptr(dm)))
        if build_pattern:
            nnz = C.c_int64()
            self._check(lib().ptb_build_pattern(self._h, C.byref(nnz)))
            self.nnz = nnz.value
        else:
            rp, cl = _a(P["rowptr"], np.int64), _a(P["cols"], np.int32)
            self._check(lib().ptb_set_pattern(self._h, _ptr(rp), _ptr(cl)))
        if device_data:
            self.locate_bc()
        else:
            bc = _a(P["bc_dofs"], np.int32)
            self._check(lib().ptb_set_bc(self._h, len(bc), _ptr(bc)))
        fc, fl = _a(P["facet_cells"], np.int32), _a(P["facet_local"], np.int32)
        self._check(lib().ptb_set_exterior_facets(self._h, len(fc), _ptr(fc), _ptr(fl)))
        if source and device_data:
            self.interpolate_source(None if P.order == 1 else P["dof_x"])
        elif source:
            self.set_source(P["f"], P["g"] if len(P["g"]) else None)
        if P.n_nbr > 0:
            self._check(lib().ptb_set_halo(
                self._h, P.n_nbr, _ptr(_a(P["nbr_ranks"], np.int32)),
                _ptr(_a(P["send_displ"], np.int32)), _ptr(_a(P["local_indices"], np.int32)),
                _ptr(_a(P["recv_displ"], np.int32)), _ptr(_a(P["remote_indices"], np.int32))))

    def set_problem_on_device(self, P):
        """The whole setup of a problem generated on the device: mesh and
        dofmap (ptb_create_box), pattern and layouts (ptb_build_pattern; with PTB_GPU_SETUP=1
        also the assembly maps), Dirichlet dofs, source terms. Only the surface-sized lists (exterior
        facets, halo) come from the host problem P, which also names the box and the rank."""
        self.P = P
        self.bs, self.nd = P.bs, P.nd
        sizes = np.zeros(4, dtype=np.int64)
        self._check(lib().ptb_create_box(self._h, PROBLEMS[P.problem_type], P.bs, P.order, P.nx, P.ny, P.nz,
                                         P.rank, P.nranks, _ptr(sizes)))
        self.n_vertices, self.n_cells, self.n_owned, self.n_ghost = (int(v) for v in sizes)
        nnz = C.c_int64()
        self._check(lib().ptb_build_pattern(self._h, C.byref(nnz)))
        self.nnz = nnz.value
        self.locate_bc()
        fc, fl = _a(P["facet_cells"], np.int32), _a(P["facet_local"], np.int32)
        self._check(lib().ptb_set_exterior_facets(self._h, len(fc), _ptr(fc), _ptr(fl)))
        self.interpolate_source(None)
        if P.n_nbr > 0:
            self._check(lib().ptb_set_halo(
                self._h, P.n_nbr, _ptr(_a(P["nbr_ranks"], np.int32)),
                _ptr(_a(P["send_displ"], np.int32)), _ptr(_a(P["local_indices"], np.int32)),
                _ptr(_a(P["recv_displ"], np.int32)), _ptr(_a(P["remote_indices"], np.int32))))

    def mesh(self, topology=True):
        """(x, x_dofmap) as the device holds them; topology=False skips the cell -> vertex map
        (returns None for it)."""
        x = np.empty(self.n_vertices * 3, dtype=np.float64)
        xd = np.empty(self.n_cells * 4, dtype=np.int32) if topology else None
        self._check(lib().ptb_get_mesh(self._h, _ptr(x), None if xd is None else _ptr(xd)))
        return x, xd

    def dofmap(self):
        dm = np.empty(self.n_cells * self.nd, dtype=np.int32)
        self._check(lib().ptb_get_dofmap(self._h, _ptr(dm)))
        return dm

    def dof_coordinates(self):
        x = np.empty((self.n_owned + self.n_ghost) * 3, dtype=np.float64)
        self._check(lib().ptb_get_dof_coordinates(self._h, _ptr(x)))
        return x

    def set_source(self, f, g=None):
        f = _a(f, np.float64)
        g = None if g is None else _a(g, np.float64)
        self._check(lib().ptb_set_source(self._h, _ptr(f), _ptr(g)))

    def locate_bc(self):
        """Dirichlet dofs located on the device (ascending local block dofs)."""
        n = C.c_int32()
        self._check(lib().ptb_locate_bc(self._h, C.byref(n)))
        out = np.empty(n.value, dtype=np.int32)
        if n.value:
            self._check(lib().ptb_get_bc(self._h, _ptr(out)))
        return out

    def interpolate_source(self, dof_x=None):
        x = None if dof_x is None else _a(dof_x, np.float64)
        self._check(lib().ptb_interpolate_source(self._h, _ptr(x)))

    def source(self):
        """(f, g | None) as the device holds them."""
        n = self.n_owned + self.n_ghost
        f = np.empty(n * self.bs, dtype=np.float64)
        g = np.empty(n, dtype=np.float64) if self.bs == 1 else None
        self._check(lib().ptb_get_source(self._h, _ptr(f), _ptr(g)))
        return f, g

    def update_geometry(self, x):
        x = _a(x, np.float64)
        self._check(lib().ptb_update_geometry(self._h, _ptr(x)))

    def comm_init(self, rank, nranks, unique_id: bytes):
        self._check(lib().ptb_comm_init(self._h, rank, nranks, C.c_char_p(unique_id)))

    def peer_export(self) -> bytes:
        buf = C.create_string_buffer(192)
        self._check(lib().ptb_peer_export(self._h, buf))
        return buf.raw

    def peer_connect(self, rank, nranks, all_handles: bytes, src_index):
        assert len(all_handles) == 192 * nranks
        si = _a(src_index, np.int32)
        self._check(lib().ptb_peer_connect(self._h, rank, nranks, C.c_char_p(all_handles),
                                           _ptr(si)))

    # ---- hot calls -----------------------------------------------------------------------
    def assemble_matrix(self):
        self._check(lib().ptb_assemble_matrix(self._h))

    def assemble_vector(self):
        self._check(lib().ptb_assemble_vector(self._h))

    def cg_solve(self, kmax=10000, rtol=1e-8, precond="jacobi"):
        it, rel = C.c_int(), C.c_double()
        self._check(lib().ptb_cg_solve(self._h, kmax, rtol, PRECOND[precond], C.byref(it),
                                       C.byref(rel)))
        return it.value, rel.value

    def set_operator_mode(self, mode):
        """'assembled' (default) or 'matrix_free' (Poisson P1: the cgpoisson action)."""
        self._check(lib().ptb_set_operator_mode(self._h, {"assembled": 0, "matrix_free": 1}[mode]))

    def set_cg_persistent(self, mode):
        """-1 auto, 0 three kernels per iteration, 1 persistent loop; across GPUs every rank must
        pass the same value (decide from the global size)."""
        self._check(lib().ptb_set_cg_persistent(self._h, int(mode)))

    def apply_operator(self, p):
        p = _a(p, np.float64)
        assert len(p) == (self.n_owned + self.n_ghost) * self.bs
        y = np.empty(self.n_owned * self.bs)
        self._check(lib().ptb_apply_operator(self._h, _ptr(p), _ptr(y)))
        return y

    # ---- data ----------------------------------------------------------------------------
    def set_rhs(self, b):
        b = _a(b, np.float64)
        assert len(b) == self.n_owned * self.bs
        self._check(lib().ptb_set_rhs(self._h, _ptr(b)))

    def set_initial_guess(self, x):
        x = None if x is None else _a(x, np.float64)
        self._check(lib().ptb_set_initial_guess(self._h, _ptr(x)))

    def matrix_values(self, out=None):
        v = np.empty(self.nnz * self.bs * self.bs) if out is None else out
        self._check(lib().ptb_get_matrix_values(self._h, _ptr(v)))
        return v

    def diagonal_inverse(self):
        v = np.empty(self.n_owned * self.bs)
        self._check(lib().ptb_get_diagonal_inverse(self._h, _ptr(v)))
        return v

    def rhs(self, out=None):
        v = np.empty(self.n_owned * self.bs) if out is None else out
        self._check(lib().ptb_get_rhs(self._h, _ptr(v)))
        return v

    def solution(self, out=None):
        v = np.empty((self.n_owned + self.n_ghost) * self.bs) if out is None else out
        self._check(lib().ptb_get_solution(self._h, _ptr(v)))
        return v

    def solution_norm(self):
        n = C.c_double()
        self._check(lib().ptb_solution_norm(self._h, C.byref(n)))
        return n.value

    def slot_offsets(self):
        n = C.c_int64()
        self._check(lib().ptb_get_slot_offsets(self._h, C.byref(n), None, None, None))
        ptr = np.empty(self.n_owned + 1, dtype=np.int64)
        pairs = np.empty(n.value, dtype=np.uint32)
        off = np.empty(n.value * self.nd, dtype=np.uint16)
        self._check(lib().ptb_get_slot_offsets(self._h, C.byref(n), _ptr(ptr), _ptr(pairs),
                                               _ptr(off)))
        return ptr, pairs, off

    def pattern(self):
        """(rowptr, cols) of a pattern built by ptb_build_pattern."""
        rp = np.empty(self.n_owned + 1, dtype=np.int64)
        self._check(lib().ptb_get_pattern(self._h, _ptr(rp), None))
        cl = np.empty(int(rp[-1]), dtype=np.int32)
        self._check(lib().ptb_get_pattern(self._h, _ptr(rp), _ptr(cl)))
        return rp, cl

    def p1_maps(self):
        """The P1 assembly maps downloaded from the device: dict(adj_off, adjrot, walk | None,
        built_on_device)."""
        ns = (self.n_owned + 31) // 32
        adj_off = np.empty(ns + 1, dtype=np.int64)
        hw, dev = C.c_int(), C.c_int()
        self._check(lib().ptb_get_p1_maps(self._h, _ptr(adj_off), None, None, C.byref(hw), C.byref(dev)))
        adjrot = np.empty(int(adj_off[-1]), dtype=np.uint32)
        walk = np.empty(int(adj_off[-1]), dtype=np.uint32) if hw.value else None
        self._check(lib().ptb_get_p1_maps(self._h, _ptr(adj_off), _ptr(adjrot), _ptr(walk), C.byref(hw),
                                          C.byref(dev)))
        return {"adj_off": adj_off, "adjrot": adjrot, "walk": walk, "built_on_device": bool(dev.value)}

    def p1_rings(self, n_sell_entries):
        """The edge rings downloaded from the device: (ring_off, ring_ns, ring) or None."""
        ns = (self.n_owned + 31) // 32
        ring_off = np.empty(ns + 1, dtype=np.int64)
        ring_ns = np.empty(n_sell_entries // 32, dtype=np.uint8)
        have = C.c_int()
        self._check(lib().ptb_get_p1_rings(self._h, _ptr(ring_off), _ptr(ring_ns), None, C.byref(have)))
        if not have.value:
            return None
        ring = np.empty(max(int(ring_off[-1]), 1), dtype=np.uint32)
        self._check(lib().ptb_get_p1_rings(self._h, None, None, _ptr(ring), C.byref(have)))
        return ring_off, ring_ns, ring

    # ---- instrumentation -------------------------------------------------------------------
    def stage_ms(self, stage):
        return lib().ptb_stage_ms(self._h, stage)

    def time_kernel(self, which, reps=20):
        """Average ms per launch of one hot kernel (KERNEL_* constants) on the resident data."""
        ms = C.c_double()
        self._check(lib().ptb_time_kernel(self._h, which, reps, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return lib().ptb_launch_count(self._h)

    def cols_explicit_fraction(self):
        return lib().ptb_cols_explicit_fraction(self._h)

    def spmv_stored_entries(self):
        return lib().ptb_spmv_stored_entries(self._h)

    def device_bytes(self):
        return lib().ptb_device_bytes(self._h)
