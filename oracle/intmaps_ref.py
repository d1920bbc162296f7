"""TEST INFRASTRUCTURE (oracle) -- numpy/pure-Python restatement of the integer side.

Independent second route for everything the product builds in C++ on the host
(performance-test_b200/host, common/intmaps.cpp): box mesh (DOLFINx create_box, tetrahedron,
called at src/mesh.cpp:184-186; Kuhn split per SURVEY B1), Lagrange dofmaps (Basix layout,
SURVEY B3), Dirichlet dof location (src/poisson_problem.cpp:58-77,
src/elasticity_problem.cpp:125-145), exterior facets, sparsity pattern (fem::create_matrix,
src/poisson_problem.cpp:122-123) and the cell -> CSR slot map. Small meshes only (Python loops).

Tests compare these arrays bit-for-bit with the product's.
"""
from __future__ import annotations

import numpy as np

KUHN = [(0, 1, 3, 7), (0, 1, 7, 5), (0, 5, 7, 4), (0, 3, 2, 7), (0, 6, 4, 7), (0, 2, 6, 7)]
EDGES = [(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)]
FACES = [(1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2)]

# Global numbering spec (DESIGN.md "Dof numbering"): level-major, each level = plane block then
# layer block; inside a block the kinds below in this order, lexicographic (iy, ix), sub-dofs
# adjacent. A kind is (dim, offsets-of-other-vertices-from-base as 3-bit codes).
PLANE_KINDS = [(0, ()), (1, (1,)), (1, (2,)), (1, (3,)), (2, (1, 3)), (2, (2, 3))]
LAYER_KINDS = [(1, (4,)), (1, (5,)), (1, (6,)), (1, (7,)),
               (2, (3, 7)), (2, (1, 7)), (2, (5, 7)), (2, (4, 7)), (2, (6, 7)), (2, (2, 7)),
               (2, (2, 6)), (2, (4, 6)), (2, (1, 5)), (2, (4, 5))]


def _bits(c):
    return np.array([c & 1, (c >> 1) & 1, (c >> 2) & 1])


def slab(nz, rank, nranks):
    base, rem = divmod(nz, nranks)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class RefProblem:
    """Same named scalars/arrays as performance-test_b200.host.Problem, built independently."""

    def __init__(self, problem_type, order, nx, ny, nz, rank=0, nranks=1):
        self.problem_type, self.order = problem_type, order
        self.nx, self.ny, self.nz, self.rank, self.nranks = nx, ny, nz, rank, nranks
        self.bs = 3 if problem_type == "elasticity" else 1
        self.nd = (order + 1) * (order + 2) * (order + 3) // 6
        L0, L1 = slab(nz, rank, nranks)
        l0 = L0 - 1 if rank > 0 else L0
        self.L0, self.L1, self.l0, self.l1 = L0, L1, l0, L1
        A = {}
        nvx, nvy = nx + 1, ny + 1

        # ---- geometry -------------------------------------------------------------------
        iz, iy, ix = np.meshgrid(np.arange(l0, L1 + 1), np.arange(nvy), np.arange(nvx),
                                 indexing="ij")
        A["x"] = np.stack([(1.0 / nx) * ix, (1.0 / ny) * iy, (1.0 / nz) * iz],
                          axis=-1).reshape(-1).astype(np.float64)
        cells_lat = []  # per local cell: 4 lattice points
        for z in range(l0, L1):
            for y in range(ny):
                for x_ in range(nx):
                    for t in KUHN:
                        cells_lat.append([np.array([x_, y, z]) + _bits(c) for c in t])
        self._cells_lat = cells_lat
        A["x_dofmap"] = np.array([[(p[2] - l0) * nvx * nvy + p[1] * nvx + p[0] for p in cell]
                                  for cell in cells_lat], dtype=np.int32).reshape(-1)
        self.n_cells = len(cells_lat)
        self.n_cells_owned = 6 * nx * ny * (L1 - L0)
        self.n_ghost_cells_front = 6 * nx * ny * (L0 - l0)
        self.cell_global_offset = 6 * nx * ny * l0
        self.n_cells_global = 6 * nx * ny * nz
        self.n_vertices = nvx * nvy * (L1 - l0 + 1)

        # ---- numbering ------------------------------------------------------------------
        def nsub(dim):
            return {0: 1, 1: order - 1, 2: (order - 1) * (order - 2) // 2}[dim]

        def block_layout(kinds):
            off, lay = 0, {}
            for (dim, offs) in kinds:
                ex = max([o & 1 for o in offs], default=0)
                ey = max([(o >> 1) & 1 for o in offs], default=0)
                w, h = nx + 1 - ex, ny + 1 - ey
                lay[(dim, offs)] = (off, w, h, nsub(dim))
                off += w * h * nsub(dim)
            return lay, off

        play, PS = block_layout(PLANE_KINDS)
        llay, LS = block_layout(LAYER_KINDS)
        self._play, self._llay, self._PS, self._LS = play, llay, PS, LS
        S = PS + LS
        total = nz * S + PS
        G0 = L0 * S
        G1 = total if rank == nranks - 1 else L1 * S
        Glow = l0 * S
        Ghigh = G1 if rank == nranks - 1 else G1 + PS
        self.n_global, self.global_offset = total, G0
        self.n_owned = G1 - G0
        n_low, n_high = G0 - Glow, Ghigh - G1
        self.n_ghost = n_low + n_high

        def to_local(g):
            if G0 <= g < G1:
                return g - G0
            if Glow <= g < G0:
                return self.n_owned + g - Glow
            if G1 <= g < Ghigh:
                return self.n_owned + n_low + g - G1
            raise AssertionError("dof outside local ranges")

        def glob(dim, offs, base, sub):
            key = (dim, tuple(offs))
            if key in play:
                off, w, h, ns = play[key]
                return base[2] * S + off + (base[1] * w + base[0]) * ns + sub
            off, w, h, ns = llay[key]
            return base[2] * S + PS + off + (base[1] * w + base[0]) * ns + sub

        self._glob, self._to_local = glob, to_local
        A["ghost_global"] = np.array(list(range(Glow, G0)) + list(range(G1, Ghigh)), dtype=np.int64)
        A["ghost_owner"] = np.array([rank - 1] * n_low + [rank + 1] * n_high, dtype=np.int32)

        # ---- dofmap + dof coordinates ---------------------------------------------------
        a_gll = 0.5 * (1 - 1 / np.sqrt(5.0))
        tpar = {2: [0.5], 3: [a_gll, 1 - a_gll]}.get(order, [])
        h = np.array([1.0 / nx, 1.0 / ny, 1.0 / nz])
        dof_x = np.zeros((self.n_owned + self.n_ghost, 3))
        dm = []
        ent_of_dof = {}
        for cell in cells_lat:
            row = []
            for p in cell:
                l = to_local(glob(0, (), p, 0))
                row.append(l)
                dof_x[l] = h * p
                ent_of_dof[l] = [tuple(p)]
            if order >= 2:
                for (a, b) in EDGES:
                    pa, pb = cell[a], cell[b]
                    lo, hi = (pa, pb) if tuple(pa[::-1]) < tuple(pb[::-1]) else (pb, pa)
                    d = hi - lo
                    code = int(d[0] + 2 * d[1] + 4 * d[2])
                    for s in range(order - 1):
                        # local dof s sits at parameter tpar[s] from local vertex a
                        gs = s if lo is pa else order - 2 - s
                        l = to_local(glob(1, (code,), lo, gs))
                        row.append(l)
                        dof_x[l] = h * (lo + tpar[gs] * d)
                        ent_of_dof[l] = [tuple(pa), tuple(pb)]
            if order == 3:
                for f in FACES:
                    pts = [cell[v] for v in f]
                    base = min(pts, key=lambda q: tuple(q[::-1]))
                    offs = sorted(int((q - base)[0] + 2 * (q - base)[1] + 4 * (q - base)[2])
                                  for q in pts if q is not base)
                    l = to_local(glob(2, tuple(offs), base, 0))
                    row.append(l)
                    dof_x[l] = h * (base + (pts[0] + pts[1] + pts[2] - 3 * base) / 3.0)
                    ent_of_dof[l] = [tuple(q) for q in pts]
            dm.append(row)
        A["dofmap"] = np.array(dm, dtype=np.int32).reshape(-1)
        A["dof_x"] = dof_x.reshape(-1)

        # ---- Dirichlet dofs: closure of facets whose vertices all satisfy the predicate ----
        if problem_type == "elasticity":
            on = lambda p: p[1] == 0
        else:
            on = lambda p: p[0] == 0 or p[0] == nx
        # an entity of the Kuhn mesh with all vertices on a coordinate plane lies in a boundary
        # facet of that plane; for x = 0 / x = 1 the two planes are distinct so check per plane
        bc = []
        for l, pts in ent_of_dof.items():
            if problem_type == "elasticity":
                ok = all(p[1] == 0 for p in pts)
            else:
                ok = all(p[0] == 0 for p in pts) or all(p[0] == nx for p in pts)
            if ok:
                bc.append(l)
        A["bc_dofs"] = np.array(sorted(bc), dtype=np.int32)
        self.n_bc = len(bc)

        # ---- RHS --------------------------------------------------------------------------
        X = dof_x
        if problem_type == "elasticity":
            dx, dz = X[:, 0] - 0.5, X[:, 2] - 0.5
            r = np.sqrt(dx * dx + dz * dz)
            A["f"] = np.stack([-dz * r * X[:, 1], np.ones(len(X)), dx * r * X[:, 1]],
                              axis=-1).reshape(-1)
            A["g"] = np.zeros(0)
        else:
            dx, dy = X[:, 0] - 0.5, X[:, 1] - 0.5
            A["f"] = 10 * np.exp(-(dx * dx + dy * dy) / 0.02)
            A["g"] = np.sin(5 * X[:, 0])

        # ---- exterior facets ----------------------------------------------------------------
        n3 = (nx, ny, nz)
        fc, fl = [], []
        for c, cell in enumerate(cells_lat):
            for lf, f in enumerate(FACES):
                pts = [cell[v] for v in f]
                for ax in range(3):
                    if all(p[ax] == 0 for p in pts) or all(p[ax] == n3[ax] for p in pts):
                        fc.append(c)
                        fl.append(lf)
        A["facet_cells"] = np.array(fc, dtype=np.int32)
        A["facet_local"] = np.array(fl, dtype=np.int32)
        self.n_facets = len(fc)

        # ---- sparsity pattern of owned rows (sorted union of cell dofs) -------------------
        dmr = A["dofmap"].reshape(-1, self.nd)
        rows = [set() for _ in range(self.n_owned)]
        for cd in dmr:
            for i in cd:
                if i < self.n_owned:
                    rows[i].update(int(j) for j in cd)
        rowptr = np.zeros(self.n_owned + 1, dtype=np.int64)
        cols = []
        for r, s in enumerate(rows):
            cs = sorted(s)
            cols.extend(cs)
            rowptr[r + 1] = rowptr[r] + len(cs)
        A["rowptr"], A["cols"] = rowptr, np.array(cols, dtype=np.int32)
        self.nnz = len(cols)

        # ---- halo lists ----------------------------------------------------------------------
        nbr, sd, rd, li, ri = [], [0], [0], [], []
        if rank > 0:
            nbr.append(rank - 1)
            li += list(range(PS))
            ri += list(range(self.n_owned, self.n_owned + n_low))
            sd.append(len(li)), rd.append(len(ri))
        if rank < nranks - 1:
            nbr.append(rank + 1)
            li += list(range(self.n_owned - S, self.n_owned))
            ri += list(range(self.n_owned + n_low, self.n_owned + n_low + n_high))
            sd.append(len(li)), rd.append(len(ri))
        for k, v in dict(nbr_ranks=nbr, send_displ=sd, recv_displ=rd, local_indices=li,
                         remote_indices=ri).items():
            A[k] = np.array(v, dtype=np.int32)
        self.n_nbr = len(nbr)
        self._A = A

    def __getitem__(self, name):
        return self._A[name]


def cell_slot_map(dofmap, nd, n_rows, rowptr, cols):
    """slot[c, i, j] = CSR position of (dofmap[c, i], dofmap[c, j]); -1 if row not owned."""
    dm = np.asarray(dofmap).reshape(-1, nd)
    slot = -np.ones((len(dm), nd, nd), dtype=np.int64)
    for c, cd in enumerate(dm):
        for i, r in enumerate(cd):
            if r >= n_rows:
                continue
            rc = cols[rowptr[r]:rowptr[r + 1]]
            for j, col in enumerate(cd):
                k = int(np.searchsorted(rc, col))
                assert rc[k] == col
                slot[c, i, j] = rowptr[r] + k
    return slot.reshape(-1)
