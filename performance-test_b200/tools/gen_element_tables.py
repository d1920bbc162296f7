"""Generates performance-test_b200/csrc/element_tables.h: exactly integrated reference tensors of
the Lagrange (basix gll_warped) P2 and P3 elements on the reference tetrahedron, used by the
CUDA element kernels in "tensor representation" (SURVEY B4):

    Ae[i][j] = sum_{b<=c} G_bc * S[bc][i][j],  G = |detJ| K K^T,
    S[bb] = int d_b phi_i d_b phi_j,   S[bc] = int (d_b phi_i d_c phi_j + d_c phi_i d_b phi_j)
    be[i]    = |detJ| sum_j M[i][j] f_j,        M  = int phi_i phi_j            (reference tet)
    facet    = |J t1 x J t2| sum_j Mf[lf][i][j] g_j, Mf = int phi_i phi_j       (reference triangle)

Route: 60-digit mpmath arithmetic, nodal basis by Vandermonde inversion in the monomial basis and
the closed form int x^a y^b z^c = a! b! c! / (a+b+c+3)!  -- independent of the oracle, which
tabulates the basis at Gauss-Jacobi points in double precision (oracle/tables.py).
Run:  python performance-test_b200/tools/gen_element_tables.py [output path]
"""
import itertools
import os

from mpmath import mp, mpf, matrix, factorial, sqrt

mp.dps = 60
EDGES = [(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)]
FACES = [(1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2)]
V = [(mpf(0), mpf(0), mpf(0)), (mpf(1), mpf(0), mpf(0)), (mpf(0), mpf(1), mpf(0)),
     (mpf(0), mpf(0), mpf(1))]


def edge_params(order):
    if order == 2:
        return [mpf(1) / 2]
    a = (1 - 1 / sqrt(mpf(5))) / 2
    return [a, 1 - a]


def nodes(order):
    pts = list(V)
    for (a, b) in EDGES:
        for t in edge_params(order):
            pts.append(tuple(V[a][d] + t * (V[b][d] - V[a][d]) for d in range(3)))
    if order == 3:
        for f in FACES:
            pts.append(tuple(sum(V[v][d] for v in f) / 3 for d in range(3)))
    return pts


def monomials(order, dim):
    return [e for e in itertools.product(range(order + 1), repeat=dim) if sum(e) <= order]


def integral(e):  # over the reference simplex of dimension len(e)
    num = mpf(1)
    for a in e:
        num *= factorial(a)
    return num / factorial(sum(e) + len(e))


def nodal_coefficients(pts, exps):
    n = len(pts)
    Vm = matrix(n, n)
    for i, p in enumerate(pts):
        for j, e in enumerate(exps):
            v = mpf(1)
            for d, a in enumerate(e):
                v *= p[d] ** a
            Vm[i, j] = v
    return Vm ** -1  # C[m, i]: phi_i = sum_m C[m, i] x^e_m


def tables(order):
    pts = nodes(order)
    nd = len(pts)
    exps = monomials(order, 3)
    C = nodal_coefficients(pts, exps)
    nm = len(exps)
    # monomial Gram matrices
    def gram(da, db):
        Gm = matrix(nm, nm)
        for m, em in enumerate(exps):
            for n, en in enumerate(exps):
                fac = mpf(1)
                e = [em[d] + en[d] for d in range(3)]
                if da is not None:
                    if em[da] == 0 or en[db] == 0:
                        continue
                    fac = mpf(em[da] * en[db])
                    e[da] -= 1
                    e[db] -= 1
                Gm[m, n] = fac * integral(e)
        return C.T * Gm * C
    M = gram(None, None)
    K = {(a, b): gram(a, b) for a in range(3) for b in range(3)}
    S = []
    for (b, c) in [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]:
        S.append(K[(b, c)] if b == c else K[(b, c)] + K[(c, b)])
    # facet mass: the trace space on facet lf is the 2-D Lagrange space on the facet's nodes
    e2 = monomials(order, 2)
    Mf = []
    for lf, f in enumerate(FACES):
        A, B, Cc = (V[v] for v in f)
        # (s, t) with X = A + s (B - A) + t (C - A): solve for each node on the facet
        on = [i for i, p in enumerate(pts) if abs(sum(p) - 1) < mpf(10) ** -40] if lf == 0 else \
             [i for i, p in enumerate(pts) if abs(p[lf - 1]) < mpf(10) ** -40]
        st = []
        for i in on:
            p = pts[i]
            # least squares on the 3x2 system (exact since p lies in the plane)
            u = [B[d] - A[d] for d in range(3)]
            w = [Cc[d] - A[d] for d in range(3)]
            r = [p[d] - A[d] for d in range(3)]
            uu, uw, ww = sum(x * x for x in u), sum(x * y for x, y in zip(u, w)), sum(x * x for x in w)
            ru, rw = sum(x * y for x, y in zip(r, u)), sum(x * y for x, y in zip(r, w))
            det = uu * ww - uw * uw
            st.append(((ru * ww - rw * uw) / det, (rw * uu - ru * uw) / det))
        assert len(on) == len(e2), (order, lf, len(on))
        C2 = nodal_coefficients(st, e2)
        G2 = matrix(len(e2), len(e2))
        for m, em in enumerate(e2):
            for n, en in enumerate(e2):
                G2[m, n] = integral([em[0] + en[0], em[1] + en[1]])
        m2 = C2.T * G2 * C2
        full = matrix(nd, nd)
        for a_, i in enumerate(on):
            for b_, j in enumerate(on):
                full[i, j] = m2[a_, b_]
        Mf.append(full)
    KF = [K[(a, b)] for a in range(3) for b in range(3)]
    return nd, S, M, Mf, KF


def emit(name, mats, nd, out):
    flat = []
    for Mx in mats:
        for i in range(nd):
            for j in range(nd):
                flat.append(Mx[i, j])
    out.append(f"static const double {name}[{len(flat)}] = {{")
    line = "  "
    for v in flat:
        s = repr(float(v)) + ", "
        if len(line) + len(s) > 98:
            out.append(line.rstrip())
            line = "  "
        line += s
    out.append(line.rstrip())
    out.append("};")


def main():
    out = ["// GENERATED by performance-test_b200/tools/gen_element_tables.py -- do not edit.",
           "// Exactly integrated reference tensors of the gll_warped Lagrange P2/P3 tetrahedron",
           "// (basix element of poisson_problem.cpp:35-38; forms Poisson.py:31-32).",
           "// S: [6][nd][nd] stiffness combos (00,01,02,11,12,22); M: [nd][nd] mass;",
           "// MF: [4][nd][nd] facet mass on the reference triangle (zero off the facet);",
           "// KF: [9][nd][nd] the unsymmetrised stiffness tensors int d_c phi_i d_d phi_j, index 3c+d",
           "// (vector-valued forms: Elasticity.py:30-39 needs d_a phi_i d_b phi_j for a != b).",
           "#pragma once", "namespace ptb { namespace tables {"]
    for order in (2, 3):
        nd, S, M, Mf, KF = tables(order)
        # sanity: rows of S sum to zero (constants in the kernel), sum(M) = 1/6, sum(Mf) = 1/2
        for Sx in S:
            for i in range(nd):
                assert abs(sum(Sx[i, j] for j in range(nd))) < mpf(10) ** -40
        assert abs(sum(M[i, j] for i in range(nd) for j in range(nd)) - mpf(1) / 6) < mpf(10) ** -40
        for Mx in Mf:
            assert abs(sum(Mx[i, j] for i in range(nd) for j in range(nd)) - mpf(1) / 2) < mpf(10) ** -40
        emit(f"S_P{order}", S, nd, out)
        emit(f"M_P{order}", [M], nd, out)
        emit(f"MF_P{order}", Mf, nd, out)
        # KF is consistent with S: S[bc] = KF[bc] + KF[cb] (b != c), S[bb] = KF[bb]
        for q, (b, c) in enumerate([(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]):
            ref = KF[3 * b + c] if b == c else KF[3 * b + c] + KF[3 * c + b]
            assert max(abs(ref[i, j] - S[q][i, j]) for i in range(nd) for j in range(nd)) < mpf(10) ** -40
        emit(f"KF_P{order}", KF, nd, out)
    out.append("} } // namespace ptb::tables")
    import sys
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(
        os.path.dirname(os.path.abspath(__file__)), "..", "csrc", "element_tables.h")
    open(path, "w").write("\n".join(out) + "\n")
    print("wrote", os.path.normpath(path))


if __name__ == "__main__":
    main()
