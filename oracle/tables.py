"""TEST INFRASTRUCTURE (oracle) -- not part of the product path.

Finite-element tables for the CPU restatement: what FFCx bakes into the generated
``tabulate_tensor`` C code of the reference (built from src/Poisson.py:15-33 and
src/Elasticity.py:11-40 at cmake time, src/CMakeLists.txt:23-40).  Un-vendored dependencies whose
published algorithm is restated here: Basix (Lagrange ``gll_warped`` element on the reference
tetrahedron, as requested at src/poisson_problem.cpp:35-38) and FFCx (quadrature degree = summed
polynomial degree; scale factor |detJ|; reference-facet scale).  parity unpinned: the reference
holds no golden vectors for this path; the tables are pinned by the analytic KATs in
tests/test_oracle_kats.py instead.

Route: numeric nodal basis (Vandermonde inversion in a monomial basis) tabulated at collapsed
Gauss-Jacobi points.  The CUDA kernels use a different route (closed forms / exactly integrated
reference tensors), so a shared mistake is unlikely.
"""
from __future__ import annotations

import itertools

import numpy as np
from scipy.special import roots_jacobi

# Basix / UFC reference tetrahedron.
REF_VERTS = np.array([[0., 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]])
TET_EDGES = [(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)]
TET_FACES = [(1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2)]


def lagrange_nodes(order: int) -> np.ndarray:
    """gll_warped Lagrange nodes in Basix dof order: vertices, edges, faces."""
    pts = [v for v in REF_VERTS]
    if order >= 2:
        if order == 2:
            ts = [0.5]
        elif order == 3:
            a = 0.5 * (1.0 - 1.0 / np.sqrt(5.0))
            ts = [a, 1.0 - a]
        else:
            raise ValueError("order must be 1..3")
        for (a_, b_) in TET_EDGES:
            for t in ts:
                pts.append(REF_VERTS[a_] + t * (REF_VERTS[b_] - REF_VERTS[a_]))
    if order == 3:
        for f in TET_FACES:
            pts.append(REF_VERTS[list(f)].mean(axis=0))
    return np.array(pts)


def _monomials(order: int):
    return [e for e in itertools.product(range(order + 1), repeat=3) if sum(e) <= order]


def _eval_monomials(exps, X, deriv=None):
    X = np.atleast_2d(X)
    out = np.empty((X.shape[0], len(exps)))
    for j, e in enumerate(exps):
        e = list(e)
        c = 1.0
        if deriv is not None:
            c = e[deriv]
            e[deriv] = max(e[deriv] - 1, 0)
        out[:, j] = c * X[:, 0] ** e[0] * X[:, 1] ** e[1] * X[:, 2] ** e[2]
    return out


class Lagrange:
    def __init__(self, order: int):
        self.order = order
        self.nodes = lagrange_nodes(order)
        self.nd = len(self.nodes)
        self.exps = _monomials(order)
        V = _eval_monomials(self.exps, self.nodes)  # V[i, j] = m_j(X_i)
        self.coef = np.linalg.inv(V)                # phi_i = sum_j coef[j, i] m_j

    def tabulate(self, X):
        """phi[q, i], dphi[q, i, d] at reference points X[q, 3]."""
        phi = _eval_monomials(self.exps, X) @ self.coef
        dphi = np.stack([_eval_monomials(self.exps, X, d) @ self.coef for d in range(3)], axis=-1)
        return phi, dphi


def _gj01(n, alpha):
    x, w = roots_jacobi(n, alpha, 0)
    return 0.5 * (x + 1.0), w / 2.0 ** (alpha + 1)


def tet_quadrature(degree: int):
    """Collapsed Gauss-Jacobi rule exact to `degree` on the reference tet (weights sum to 1/6)."""
    n = max(1, (degree + 2) // 2)
    r, wr = _gj01(n, 2)
    s, ws = _gj01(n, 1)
    t, wt = _gj01(n, 0)
    pts, wts = [], []
    for i in range(n):
        for j in range(n):
            for k in range(n):
                pts.append([r[i], s[j] * (1 - r[i]), t[k] * (1 - r[i]) * (1 - s[j])])
                wts.append(wr[i] * ws[j] * wt[k])
    return np.array(pts), np.array(wts)


def tri_quadrature(degree: int):
    """Collapsed rule on the reference triangle (weights sum to 1/2)."""
    n = max(1, (degree + 2) // 2)
    r, wr = _gj01(n, 1)
    s, ws = _gj01(n, 0)
    pts, wts = [], []
    for i in range(n):
        for j in range(n):
            pts.append([r[i], s[j] * (1 - r[i])])
            wts.append(wr[i] * ws[j])
    return np.array(pts), np.array(wts)


def facet_points(lf: int, P2d: np.ndarray) -> np.ndarray:
    """Map reference-triangle points onto reference-tet facet lf (opposite vertex lf)."""
    a, b, c = (REF_VERTS[v] for v in TET_FACES[lf])
    return a[None, :] + P2d[:, :1] * (b - a)[None, :] + P2d[:, 1:2] * (c - a)[None, :]


def element_tables(order: int) -> dict:
    """All tables the C oracle needs for one polynomial order (FFCx quadrature degrees, B4)."""
    el = Lagrange(order)
    el1 = Lagrange(1)
    qa, wa = tet_quadrature(2 * (order - 1))  # a, M: grad.grad
    ql, wl = tet_quadrature(2 * order)        # L cells: f * v
    qf2, wf = tri_quadrature(2 * order)       # L facets: g * v
    _, dphi_a = el.tabulate(qa)
    phi_l, _ = el.tabulate(ql)
    phi_f = np.stack([el.tabulate(facet_points(lf, qf2))[0] for lf in range(4)])
    _, dgeo = el1.tabulate(np.zeros((1, 3)))  # affine geometry: constant gradients
    # Reference-facet edge vectors t1, t2 (facet scale = |J t1 x J t2|).
    ft = np.array([[REF_VERTS[f[1]] - REF_VERTS[f[0]], REF_VERTS[f[2]] - REF_VERTS[f[0]]]
                   for f in TET_FACES])
    c = np.ascontiguousarray
    return dict(order=order, nd=el.nd, nodes=el.nodes,
                nq_a=len(wa), w_a=c(wa), dphi_a=c(dphi_a),
                nq_l=len(wl), w_l=c(wl), phi_l=c(phi_l),
                nq_f=len(wf), w_f=c(wf), phi_f=c(phi_f),
                dgeo=c(dgeo[0]), facet_t=c(ft))
