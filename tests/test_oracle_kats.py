"""Pins the CPU oracle with the analytic known-answer tests of SURVEY 4.3 (the reference has no
golden vectors: 'parity unpinned')."""
import numpy as np
import pytest
import scipy.sparse as sp


def _csr(P, v):
    if P.bs == 1:
        return sp.csr_matrix((v, P["cols"], P["rowptr"]), shape=(P.n_owned, P.n_owned + P.n_ghost))
    return sp.bsr_matrix((v.reshape(-1, 3, 3), P["cols"], P["rowptr"]),
                         shape=(3 * P.n_owned, 3 * (P.n_owned + P.n_ghost))).tocsr()


def _single_tet(oracle, ptype, order, X):
    """A one-cell 'problem' on an arbitrary tetrahedron X[4,3]."""
    from oracle import tables

    class One:
        pass
    P = One()
    nd = (order + 1) * (order + 2) * (order + 3) // 6
    P.problem_type, P.order, P.bs, P.nd = ptype, order, 3 if ptype == "elasticity" else 1, nd
    P.n_cells, P.n_owned, P.n_ghost = 1, nd, 0
    A = dict(x=np.asarray(X, float).reshape(-1), x_dofmap=np.arange(4, dtype=np.int32),
             dofmap=np.arange(nd, dtype=np.int32),
             rowptr=np.arange(0, nd * nd + 1, nd, dtype=np.int64),
             cols=np.tile(np.arange(nd, dtype=np.int32), nd), bc_dofs=np.zeros(0, np.int32),
             f=np.ones(nd * P.bs), g=np.zeros(0), facet_cells=np.zeros(0, np.int32),
             facet_local=np.zeros(0, np.int32))
    P.__class__.__getitem__ = lambda self, k: A[k]
    return P, A


REF = np.array([[0., 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]])


def test_K1_p1_poisson_reference_tet(oracle):
    P, _ = _single_tet(oracle, "poisson", 1, REF)
    Ae = oracle.assemble_matrix(P).reshape(4, 4)
    ref = np.array([[3, -1, -1, -1], [-1, 1, 0, 0], [-1, 0, 1, 0], [-1, 0, 0, 1]]) / 6.0
    np.testing.assert_allclose(Ae, ref, atol=1e-15)


def test_K2_p1_mass_any_tet(oracle):
    rng = np.random.default_rng(0)
    X = REF + 0.2 * rng.standard_normal((4, 3))
    P, A = _single_tet(oracle, "poisson", 1, X)
    f = rng.standard_normal(4)
    A["f"] = f
    be = oracle.assemble_vector(P)
    det = abs(np.linalg.det((X[1:] - X[0]).T))
    np.testing.assert_allclose(be, det / 120.0 * (f.sum() + f), rtol=1e-13)


def test_K3_seven_point_stencil(pt, oracle, nobc):
    i, j, k = 4, 3, 5
    P = pt.host.Problem("poisson", 1, i, j, k)
    A = _csr(P, oracle.assemble_matrix(nobc(P))).toarray()
    hx, hy, hz = 1 / i, 1 / j, 1 / k
    r = (2 * (j + 1) + 1) * (i + 1) + 2  # interior vertex (ix, iy, iz) = (2, 1, 2)
    assert A[r, r] == pytest.approx(2 * (hy * hz / hx + hx * hz / hy + hx * hy / hz), rel=1e-13)
    assert A[r, r + 1] == pytest.approx(-hy * hz / hx, rel=1e-13)
    assert A[r, r + (i + 1)] == pytest.approx(-hx * hz / hy, rel=1e-13)
    assert A[r, r + (i + 1) * (j + 1)] == pytest.approx(-hx * hy / hz, rel=1e-13)
    row = A[r].copy()
    for o in (0, 1, -1, i + 1, -(i + 1), (i + 1) * (j + 1), -(i + 1) * (j + 1)):
        row[r + o] = 0
    assert abs(row).max() < 1e-15  # the 8 diagonal-edge entries are analytic zeros (D9)


@pytest.mark.parametrize("order", [1, 2, 3])
def test_K4_poisson_constants_in_kernel_and_symmetry(pt, oracle, nobc, order):
    P = pt.host.Problem("poisson", order, 3, 2, 4)
    A = _csr(P, oracle.assemble_matrix(nobc(P)))
    assert abs(A @ np.ones(P.n_owned)).max() < 1e-13
    assert abs(A - A.T).max() < 1e-14


@pytest.mark.parametrize("order", [1, 2])
def test_K4_elasticity_rigid_body_modes(pt, oracle, nobc, order):
    P = pt.host.Problem("elasticity", order, 3, 2, 3)
    A = _csr(P, oracle.assemble_matrix(nobc(P)))
    X = P["dof_x"].reshape(-1, 3)
    modes = []
    for c in range(3):
        m = np.zeros_like(X); m[:, c] = 1; modes.append(m)
    for (a, b) in ((0, 1), (2, 0), (1, 2)):  # elasticity_problem.cpp:63-70
        m = np.zeros_like(X); m[:, a] = -X[:, b]; m[:, b] = X[:, a]; modes.append(m)
    for m in modes:
        assert abs(A @ m.reshape(-1)).max() / abs(A).max() < 1e-13
    assert abs(A - A.T).max() / abs(A).max() < 1e-14


def test_K5_elasticity_p1_closed_form(oracle):
    rng = np.random.default_rng(1)
    X = REF + 0.2 * rng.standard_normal((4, 3))
    P, _ = _single_tet(oracle, "elasticity", 1, X)
    Ae = _csr(P, oracle.assemble_matrix(P)).toarray()
    J = (X[1:] - X[0]).T
    K = np.linalg.inv(J)
    g = np.vstack([-K.sum(axis=0), K])  # grad phi_i = K^T ghat_i
    mu, lm = 384615.3846153846, 576923.0769230769  # Elasticity.py:12-15 as Python doubles
    det = abs(np.linalg.det(J))
    ref = np.zeros((12, 12))
    for i in range(4):
        for a in range(3):
            for j in range(4):
                for b in range(3):
                    ref[3 * i + a, 3 * j + b] = det / 6 * (
                        mu * ((a == b) * g[i] @ g[j] + g[i][b] * g[j][a]) + lm * g[i][a] * g[j][b])
    np.testing.assert_allclose(Ae, ref, rtol=1e-12, atol=1e-9 * abs(ref).max())


@pytest.mark.parametrize("order", [2, 3])
def test_K6_high_order_reference_matrices_exact(oracle, order):
    """P2/P3 stiffness and mass on the reference tet against exact sympy integration of the nodal
    basis on the gll_warped node set."""
    import sympy as s
    from oracle import tables
    x, y, z = s.symbols("x y z")
    nodes = tables.lagrange_nodes(order)
    mons = [x**a * y**b * z**c for (a, b, c) in tables._monomials(order)]
    a_gll = (1 - 1 / s.sqrt(5)) / 2
    def exact(v):  # snap node coordinates to exact values
        for cand in (s.Integer(0), s.Integer(1), s.Rational(1, 2), s.Rational(1, 3), a_gll, 1 - a_gll):
            if abs(float(cand) - v) < 1e-12:
                return cand
        raise AssertionError(v)
    N = [[exact(v) for v in p] for p in nodes]
    V = s.Matrix([[m.subs({x: p[0], y: p[1], z: p[2]}) for m in mons] for p in N])
    C = V.inv()
    phi = [sum(C[j, i] * mons[j] for j in range(len(mons))) for i in range(len(mons))]
    def integ(e):
        return s.integrate(s.integrate(s.integrate(e, (z, 0, 1 - x - y)), (y, 0, 1 - x)), (x, 0, 1))
    nd = len(phi)
    idx = [(0, 0), (0, 1), (1, nd - 1), (4, 5), (nd - 1, nd - 1), (2, 7)]
    P, A = _single_tet(oracle, "poisson", order, REF)
    Ae = oracle.assemble_matrix(P).reshape(nd, nd)
    A["f"] = np.zeros(nd); A["f"][3] = 1.0
    be = oracle.assemble_vector(P)
    for (i, j) in idx:
        kij = integ(sum(s.diff(phi[i], v) * s.diff(phi[j], v) for v in (x, y, z)))
        assert Ae[i, j] == pytest.approx(float(kij), rel=1e-11, abs=1e-14)
    for i in (0, 3, 5, nd - 1):
        assert be[i] == pytest.approx(float(integ(phi[i] * phi[3])), rel=1e-11, abs=1e-15)


def test_K7_orientation_invariance(oracle):
    rng = np.random.default_rng(2)
    X = REF + 0.1 * rng.standard_normal((4, 3))
    Xm = X.copy(); Xm[:, 0] *= -1  # mirror image: detJ changes sign
    for ptype in ("poisson", "elasticity"):
        P, _ = _single_tet(oracle, ptype, 2, X)
        Q, _ = _single_tet(oracle, ptype, 2, Xm)
        a, b = oracle.assemble_matrix(P), oracle.assemble_matrix(Q)
        if ptype == "poisson":
            np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-13)
        else:  # mirrored x-component flips sign of the (x, y|z) couplings
            S = np.tile([-1.0, 1, 1], 10)
            np.testing.assert_allclose(_csr(P, a).toarray(),
                                       S[:, None] * _csr(Q, b).toarray() * S[None, :],
                                       rtol=1e-11, atol=1e-6)


@pytest.mark.parametrize("order", [1, 2, 3])
def test_K8_rhs_sums_to_integrals_of_interpolants(pt, oracle, nobc, order):
    """sum_i b_i = int f_h dx + int g_h ds (partition of unity), by independent quadrature."""
    from oracle import tables
    P = pt.host.Problem("poisson", order, 3, 2, 2)
    b = oracle.assemble_vector(nobc(P))
    el = tables.Lagrange(order)
    q3, w3 = tables.tet_quadrature(6)
    q2, w2 = tables.tri_quadrature(6)
    phi3 = el.tabulate(q3)[0]
    X = P["x"].reshape(-1, 3); xd = P["x_dofmap"].reshape(-1, 4); dm = P["dofmap"].reshape(-1, P.nd)
    total = 0.0
    for c in range(P.n_cells):
        det = abs(np.linalg.det((X[xd[c, 1:]] - X[xd[c, 0]]).T))
        total += det * (w3 @ (phi3 @ P["f"][dm[c]]))
    for c, lf in zip(P["facet_cells"], P["facet_local"]):
        v = [X[xd[c, t]] for t in tables.TET_FACES[lf]]
        area2 = np.linalg.norm(np.cross(v[1] - v[0], v[2] - v[0]))
        phif = el.tabulate(tables.facet_points(lf, q2))[0]
        total += area2 * (w2 @ (phif @ P["g"][dm[c]]))
    assert b.sum() == pytest.approx(total, rel=1e-12)


def test_K9_manufactured_solution_converges(pt, oracle):
    """-Laplace u = 2, u = x(1-x) (zero flux on the y/z faces): on this mesh the P1 operator is
    the 7-point stencil and the load of a constant is h^3 per interior node, so the discrete
    solution is nodally exact; CG iteration counts grow with 1/h."""
    errs, its = [], []
    for n in (4, 8, 16):
        P = pt.host.Problem("poisson", 1, n, n, n)

        class Q:
            def __getattr__(s, k): return getattr(P, k)
            def __getitem__(s, k):
                if k == "f": return 2.0 * np.ones(P.n_owned)
                if k == "g": return np.zeros(0)
                return P[k]
        A = oracle.assemble_matrix(P)
        b = oracle.assemble_vector(Q())
        x, k, rel = oracle.cg(1, P.n_owned, P["rowptr"], P["cols"], A, b, kmax=2000, rtol=1e-10,
                              precond="jacobi")
        X = P["dof_x"].reshape(-1, 3)[:, 0]
        errs.append(np.sqrt(np.mean((x - X * (1 - X)) ** 2)))
        its.append(k)
        assert rel < 1e-10
    assert max(errs) < 1e-9
    assert its[0] < its[1] < its[2]


def test_cg_none_reproduces_cg_h_semantics(pt, oracle):
    """kmax cap, >= 1 iteration, returned k, stopping on |r|^2/|r0|^2 < rtol^2 (cg.h:53-85)."""
    P = pt.host.Problem("poisson", 1, 6, 6, 6)
    A, b = oracle.assemble_matrix(P), oracle.assemble_vector(P)
    args = (1, P.n_owned, P["rowptr"], P["cols"], A, b)
    x, k, rel = oracle.cg(*args, kmax=5, rtol=1e-30)
    assert k == 5
    x, k, rel = oracle.cg(*args, kmax=50, rtol=1e3)
    assert k == 1
    x, k, rel = oracle.cg(*args, kmax=500, rtol=1e-8)
    M = _csr(P, A)
    assert rel < 1e-8 and np.linalg.norm(M @ x - b) / np.linalg.norm(b) < 2e-8
    # textbook CG in numpy, same loop
    xr = np.zeros_like(b); r = b - M @ xr; p = r.copy(); rn0 = rn = r @ r; kk = 0
    while kk < 500:
        kk += 1; y = M @ p; al = rn / (p @ y); xr += al * p; r -= al * y
        rnn = r @ r; be = rnn / rn; rn = rnn
        if rn / rn0 < 1e-16: break
        p = be * p + r
    assert abs(kk - k) <= 1
    np.testing.assert_allclose(x, xr, rtol=1e-6, atol=1e-10)


@pytest.mark.parametrize("order,rate", [(1, 1.7), (2, 2.7), (3, 3.6)])
def test_manufactured_solution_convergence_rates(pt, oracle, order, rate):
    """-Laplace u = 2 pi^2 u with u = sin(pi x) cos(pi y): u = 0 on x = 0, 1 (the reference's
    Dirichlet set) and zero flux on the other faces, so the reference's problem setup applies
    unchanged with f = 2 pi^2 u, g = 0. The nodal error must fall like h^(k+1): this pins the
    global consistency of the P2/P3 dofmaps (edge orientation, gll_warped points, face dofs) --
    a wrong edge flip in one tet type would destroy the rate."""
    errs = []
    sizes = {1: (6, 12), 2: (3, 6), 3: (2, 4)}[order]
    for n in sizes:
        P = pt.host.Problem("poisson", order, n, n, n)
        X = P["dof_x"].reshape(-1, 3)
        u_ex = np.sin(np.pi * X[:, 0]) * np.cos(np.pi * X[:, 1])

        class Q:
            def __getattr__(s, k): return getattr(P, k)
            def __getitem__(s, k):
                if k == "f": return 2 * np.pi ** 2 * u_ex
                if k == "g": return np.zeros(0)
                return P[k]
        A = oracle.assemble_matrix(P)
        b = oracle.assemble_vector(Q())
        x, k, rel = oracle.cg(1, P.n_owned, P["rowptr"], P["cols"], A, b, kmax=5000, rtol=1e-12,
                              precond="jacobi")
        assert rel < 1e-12
        errs.append(np.sqrt(np.mean((x - u_ex) ** 2)))
    observed = np.log2(errs[0] / errs[1])
    assert observed > rate, (errs, observed)


@pytest.mark.parametrize("ptype,order", [("poisson", 1), ("poisson", 2), ("poisson", 3), ("elasticity", 1)])
def test_patch_test_on_a_jittered_mesh(pt, oracle, nobc, perturbed, ptype, order):
    """On a mesh with moved vertices (general tetrahedra): the matrix stays symmetric, and a linear
    field is reproduced exactly -- its operator image vanishes on every row of a dof strictly
    inside the cube (constant gradient / constant stress has zero divergence). The Pk nodes of
    the moved mesh are the affine images of the lattice nodes (cell by cell), whatever their
    placement on the reference element."""
    dims = (4, 3, 5) if order == 1 else (3, 2, 3)
    P0 = pt.host.Problem(ptype, order, *dims)
    P = perturbed(nobc(P0))
    A = _csr(P0, oracle.assemble_matrix(P))
    assert abs(A - A.T).max() <= 1e-12 * abs(A).max()
    # physical coordinates of every dof: affine map of each cell applied to the reference nodes
    X = np.array(P["x"]).reshape(-1, 3)
    xd = np.array(P0["x_dofmap"]).reshape(-1, 4)
    dm = np.array(P0["dofmap"]).reshape(-1, P0.nd)
    X0 = np.array(P0["x"]).reshape(-1, 3)          # lattice vertices
    lat = np.array(P0["dof_x"]).reshape(-1, 3)     # lattice position of every dof
    xdof = np.zeros((P0.n_owned + P0.n_ghost, 3))
    for c in range(dm.shape[0]):
        V0, V = X0[xd[c]], X[xd[c]]
        xi = np.linalg.solve((V0[1:] - V0[0]).T, (lat[dm[c]] - V0[0]).T).T  # reference coordinates
        xdof[dm[c]] = V[0] + xi @ (V[1:] - V[0])
    a = np.array([0.3, -1.1, 0.7])
    if ptype == "poisson":
        u = xdof @ a + 0.25
    else:
        M = np.array([[0.2, -0.4, 0.1], [0.5, 0.3, -0.2], [-0.6, 0.1, 0.4]])
        u = (xdof @ M.T + np.array([0.1, -0.2, 0.3])).reshape(-1)
    r = A @ u
    interior = np.all((lat > 1e-9) & (lat < 1 - 1e-9), axis=1)
    # phi_i vanishes on the boundary for every dof strictly inside the cube: no flux term there
    rows = np.repeat(interior[:P0.n_owned], P0.bs)
    assert rows.any()
    assert np.abs(r[rows]).max() <= 1e-10 * np.abs(r).max()


@pytest.mark.parametrize("ptype,dims", [("poisson", (7, 6, 8)), ("elasticity", (4, 5, 3))])
def test_jacobi_cg_iterates_equal_scipys_preconditioned_cg(pt, oracle, ptype, dims):
    """cg.h has no preconditioner; the Jacobi extension of the oracle (z = D^-1 r, alpha = r.z / p.Ap,
    beta = r'.z' / r.z) is the textbook preconditioned CG. Pin it to an independent, widely
    used implementation: after k iterations the oracle's x equals the k-th iterate of
    scipy.sparse.linalg.cg with M = D^-1 (same recurrences, different code) to 1e-10, for several k,
    and the converged solutions agree with a sparse direct solve."""
    import scipy.sparse.linalg as spla
    P = pt.host.Problem(ptype, 1, *dims)
    bs, n = P.bs, P.n_owned * P.bs
    vals, b = oracle.assemble_matrix(P), oracle.assemble_vector(P)
    A = sp.csr_matrix(_csr(P, vals))[:, :n]
    dinv = 1.0 / A.diagonal()
    M = spla.LinearOperator((n, n), matvec=lambda r: dinv * r)
    for k in (1, 2, 5, 17):
        x_o, k_o, _ = oracle.cg(bs, P.n_owned, P["rowptr"], P["cols"], vals, b, kmax=k, rtol=1e-30, precond="jacobi")
        iterates = []
        spla.cg(A, b, x0=np.zeros(n), rtol=0.0, atol=0.0, maxiter=k, M=M, callback=lambda xk: iterates.append(xk.copy()))
        assert k_o == k and len(iterates) == k
        assert np.abs(x_o - iterates[-1]).max() <= 1e-10 * np.abs(iterates[-1]).max()
    x_o, k_o, rel = oracle.cg(bs, P.n_owned, P["rowptr"], P["cols"], vals, b, kmax=5000, rtol=1e-10, precond="jacobi")
    x_d = spla.spsolve(sp.csc_matrix(A), b)
    assert rel < 1e-10 and np.abs(x_o - x_d).max() <= 1e-7 * np.abs(x_d).max()
