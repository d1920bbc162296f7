"""Dof renumbering of the host stand-in (pth_problem_renumber): a real DOLFINx dofmap is not numbered
lattice-lexicographically (graph reordering), so the path is also exercised on "rcm" and
"random" numberings. CPU side: the renumbered problem is the same problem (oracle solution equal
after matching dofs by their coordinates), the pattern is the permuted pattern, rcm is banded."""
import numpy as np
import pytest

CASES = [("poisson", 1, (6, 5, 7)), ("elasticity", 1, (4, 5, 3)), ("poisson", 2, (3, 4, 3)), ("poisson", 3, (2, 3, 2))]


def _by_coordinate(P):
    return np.lexsort(np.round(P["dof_x"].reshape(-1, 3) * 1e7).astype(np.int64).T)


@pytest.mark.parametrize("kind", ["rcm", "random"])
@pytest.mark.parametrize("ptype,order,dims", CASES)
def test_renumbered_problem_is_the_same_problem(pt, oracle, ptype, order, dims, kind):
    P0 = pt.host.Problem(ptype, order, *dims)
    P = pt.host.Problem(ptype, order, *dims, renumber=kind, seed=3)
    assert (P.n_owned, P.nnz, P.n_bc) == (P0.n_owned, P0.nnz, P0.n_bc)
    o0, o = _by_coordinate(P0), _by_coordinate(P)
    new_of_old = np.empty(P.n_owned, dtype=np.int64)
    new_of_old[o0] = o
    assert not np.array_equal(new_of_old, np.arange(P.n_owned))
    bs = P.bs
    # every dof-indexed array followed the permutation
    assert np.array_equal(P["dof_x"].reshape(-1, 3)[new_of_old], P0["dof_x"].reshape(-1, 3))
    assert np.array_equal(P["f"].reshape(-1, bs)[new_of_old], P0["f"].reshape(-1, bs))
    assert np.array_equal(np.sort(new_of_old[P0["bc_dofs"]]), P["bc_dofs"])
    assert np.array_equal(new_of_old[P0["dofmap"]], P["dofmap"])
    # the pattern is the permuted pattern, columns ascending
    rp0, cl0, rp, cl = P0["rowptr"], P0["cols"], P["rowptr"], P["cols"]
    for r0 in range(0, P.n_owned, max(1, P.n_owned // 50)):
        r = new_of_old[r0]
        assert np.array_equal(np.sort(new_of_old[cl0[rp0[r0]:rp0[r0 + 1]]]), cl[rp[r]:rp[r + 1]])
    # and the solution is the permuted solution
    A0, b0 = oracle.assemble_matrix(P0), oracle.assemble_vector(P0)
    A, b = oracle.assemble_matrix(P), oracle.assemble_vector(P)
    x0, k0, _ = oracle.cg(bs, P0.n_owned, rp0, cl0, A0, b0, kmax=5000, rtol=1e-10, precond="jacobi")
    x, k, _ = oracle.cg(bs, P.n_owned, rp, cl, A, b, kmax=5000, rtol=1e-10, precond="jacobi")
    assert abs(k - k0) <= 1
    assert np.abs(x.reshape(-1, bs)[new_of_old] - x0.reshape(-1, bs)).max() <= 1e-9 * np.abs(x0).max()


def test_rcm_is_banded_and_random_is_not(pt):
    P0 = pt.host.Problem("poisson", 1, 12, 11, 13)
    bw = {}
    for kind in (None, "rcm", "random"):
        P = pt.host.Problem("poisson", 1, 12, 11, 13, renumber=kind, seed=5)
        rows = np.repeat(np.arange(P.n_owned), np.diff(P["rowptr"]))
        bw[kind] = int(np.abs(P["cols"] - rows).max())
    assert bw["rcm"] <= 1.5 * bw[None] and bw["random"] > 4 * bw[None]


def test_renumbering_is_refused_on_partitions(pt):
    with pytest.raises(RuntimeError, match="single-rank"):
        pt.host.Problem("poisson", 1, 4, 4, 6, 0, 2, renumber="rcm")
