"""Integer side: sizing replay, mesh/dofmap/sparsity/slot map -- product C++ vs the independent
numpy restatement in oracle/intmaps_ref.py (bit-identical), plus closed-form counts."""
import json
import os

import numpy as np
import pytest

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sizing.json")))


@pytest.mark.parametrize("case", GOLD, ids=[c["name"] for c in GOLD])
def test_sizing_matches_golden(pt, case):
    """create_cube_mesh sizing, src/mesh.cpp:78-151, against tests/golden/sizing.json."""
    got = pt.host.cube_sizing(case["target"], case["total"], case["dofs_per_node"], case["order"],
                              case["nproc"])
    assert list(got) == case["sizing"]
    Nx, Ny, Nz, r = got
    assert list(pt.host.num_entities(Nx, Ny, Nz, r)) == case["entities"]
    assert pt.host.num_pdofs(Nx, Ny, Nz, r, case["order"]) == case["pdofs"]


def test_order_not_supported(pt):
    with pytest.raises(RuntimeError, match="Order not supported"):
        pt.host.num_pdofs(2, 2, 2, 0, 5)
    with pytest.raises(RuntimeError, match="Unknown problem type"):
        pt.host.Problem("stokes", 1, 2, 2, 2)


CASES = [("poisson", 1, (2, 3, 4), 1), ("poisson", 2, (3, 2, 3), 1), ("poisson", 3, (2, 2, 3), 1),
         ("elasticity", 1, (3, 2, 5), 2), ("poisson", 1, (2, 2, 7), 3), ("poisson", 2, (2, 3, 4), 2),
         ("elasticity", 3, (2, 2, 4), 2), ("poisson", 3, (1, 1, 1), 1)]


@pytest.mark.parametrize("ptype,order,dims,nranks", CASES)
def test_host_arrays_bit_identical_to_numpy_restatement(pt, ptype, order, dims, nranks):
    from oracle import intmaps_ref as R
    for rank in range(nranks):
        P = pt.host.Problem(ptype, order, *dims, rank, nranks)
        Q = R.RefProblem(ptype, order, *dims, rank, nranks)
        for s in pt.host.SCALARS:
            assert getattr(P, s) == getattr(Q, s), s
        for a in pt.host.ARRAYS:
            x, y = P[a], Q[a]
            assert x.shape == y.shape, a
            if x.dtype.kind == "f":
                np.testing.assert_allclose(x, y, rtol=1e-14, atol=1e-15, err_msg=a)
            else:
                assert np.array_equal(x, y), a


@pytest.mark.parametrize("order", [1, 2, 3])
@pytest.mark.parametrize("dims", [(4, 3, 5), (2, 6, 3)])
def test_counts_match_closed_forms(pt, order, dims):
    """DOFs = mesh.cpp:56-74; nnz = SURVEY 8d closed forms; facets = 4(ij+ik+jk)."""
    i, j, k = dims
    P = pt.host.Problem("poisson", order, i, j, k)
    s, t = i * j + i * k + j * k, i + j + k
    assert P.n_global == pt.host.num_pdofs(i, j, k, 0, order) == P.n_owned
    nnz = {1: 15 * i * j * k + 7 * s + 3 * t + 1, 2: 230 * i * j * k + 46 * s + 8 * t + 1,
           3: 1311 * i * j * k + 153 * s + 15 * t + 1}[order]
    assert P.nnz == nnz
    assert P.n_facets == 4 * s
    assert P.n_cells == 6 * i * j * k


@pytest.mark.parametrize("ptype,order,dims,nranks", CASES[:6])
def test_cell_slot_map_bit_identical(pt, ptype, order, dims, nranks):
    """The cell -> CSR-slot map (host-only entry point of the C-ABI) vs the numpy restatement."""
    from oracle import intmaps_ref as R
    for rank in range(nranks):
        P = pt.host.Problem(ptype, order, *dims, rank, nranks)
        got = pt.abi.build_cell_slot_map(P["dofmap"], P.nd, P.n_owned, P["rowptr"], P["cols"])
        ref = R.cell_slot_map(P["dofmap"], P.nd, P.n_owned, P["rowptr"], P["cols"])
        assert np.array_equal(got, ref)


def test_partition_covers_global_problem(pt):
    """Owned ranges tile the global numbering; every rank's owned rows equal the serial rows."""
    dims, order = (3, 2, 7), 2
    S = pt.host.Problem("poisson", order, *dims)
    cols_g = [None] * S.n_owned
    tot = 0
    for rank in range(3):
        P = pt.host.Problem("poisson", order, *dims, rank, 3)
        assert P.global_offset == tot
        tot += P.n_owned
        l2g = np.concatenate([np.arange(P.global_offset, P.global_offset + P.n_owned),
                              P["ghost_global"]])
        rp, cl = P["rowptr"], P["cols"]
        for r in range(P.n_owned):
            got = np.sort(l2g[cl[rp[r]:rp[r + 1]]])
            ref = S["cols"][S["rowptr"][P.global_offset + r]:S["rowptr"][P.global_offset + r + 1]]
            assert np.array_equal(got, ref)
        np.testing.assert_array_equal(P["dof_x"].reshape(-1, 3)[:P.n_owned],
                                      S["dof_x"].reshape(-1, 3)[P.global_offset:tot])
    assert tot == S.n_global


@pytest.mark.parametrize("order,dims,nranks", [(1, (40, 37, 9), 1), (1, (5, 4, 6), 2), (2, (7, 6, 5), 1),
                                               (3, (4, 3, 5), 2), (1, (1, 1, 1), 1)])
def test_device_layout_roundtrip(pt, order, dims, nranks):
    """SELL-32 + warp-uniform column-delta compression decodes to the original CSR columns."""
    for rank in range(nranks):
        P = pt.host.Problem("poisson", order, *dims, rank, nranks)
        got, frac = pt.abi.layout_roundtrip(P.n_owned, P.n_owned + P.n_ghost, P["rowptr"], P["cols"])
        assert np.array_equal(got, P["cols"])
        assert 0.0 <= frac <= 1.0
    if dims == (40, 37, 9):
        assert frac < 0.6  # translation-invariant interior: most indices collapse to deltas


def test_abi_library_exports_every_declared_symbol(pt):
    """-m "not gpu": the C-ABI library loads and exports all of include/ptb200.h; no compute."""
    L = pt.abi.lib()
    names = pt.abi.declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), n


def test_no_cpu_fallback(pt):
    """Without a CUDA device the product must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        pt.abi.Context(0)


@pytest.mark.parametrize("ptype,order,dims,rank,nranks", [("poisson", 1, (5, 4, 6), 0, 1), ("poisson", 1, (4, 3, 5), 1, 2),
                                                          ("poisson", 2, (3, 2, 4), 0, 1), ("poisson", 3, (2, 3, 2), 1, 2),
                                                          ("elasticity", 1, (3, 4, 3), 0, 1)])
def test_facet_rows_equal_a_numpy_restatement(pt, ptype, order, dims, rank, nranks):
    """Boundary-facet gather lists of assemble_vector (layout.cpp build_facet_rows): for every owned
    row the (cell, local_facet*nd + li) entries of the exterior facets that contain the dof, rows
    ascending, entries in facet order. Closure of a facet in the Basix layout: vertices, edges, faces."""
    import numpy as np
    P = pt.host.Problem(ptype, order, *dims, rank, nranks)
    nd, ne, nf = P.nd, order - 1, (order - 1) * (order - 2) // 2
    edges = [(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)]
    on = []
    for lf in range(4):
        dofs = [v for v in range(4) if v != lf]
        dofs += [4 + e * ne + s for e, (a, b) in enumerate(edges) if a != lf and b != lf for s in range(ne)]
        dofs += [4 + 6 * ne + lf * nf + s for s in range(nf)]
        on.append(dofs)
    dm = np.asarray(P["dofmap"]).reshape(-1, nd)
    rows = {}
    for c, lf in zip(P["facet_cells"], P["facet_local"]):
        for li in on[lf]:
            r = int(dm[c, li])
            if r < P.n_owned:
                rows.setdefault(r, []).extend([int(c), int(lf) * nd + li])
    ids_ref = np.array(sorted(rows), np.int32)
    ent_ref = np.array([v for r in sorted(rows) for v in rows[r]], np.int32)
    ptr_ref = np.concatenate([[0], np.cumsum([len(rows[r]) // 2 for r in sorted(rows)])]).astype(np.int32)
    ids, ptr, ent = pt.abi.facet_rows(P["facet_cells"], P["facet_local"], P["dofmap"], nd, order, P.n_owned)
    assert np.array_equal(ids, ids_ref) and np.array_equal(ptr, ptr_ref) and np.array_equal(ent, ent_ref)


def test_host_sizes_only_modes_agree_with_the_full_build(tmp_path):
    """create_box_mesh(..., with_arrays=false) and create_functionspace(..., with_dofmap=false) --
    what the CLI keeps on the host under --device_setup -- give the same sizes, ranges, ghost, halo
    and exterior-facet lists as the full build, on every rank of a partition, P1-P3."""
    import ctypes
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    out = str(tmp_path / "libhostmodes.so")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-fopenmp", "-w", "-o", out,
                    os.path.join(here, "emu", "host_header_modes.cpp")], check=True)
    lib = ctypes.CDLL(out)
    lib.header_modes_agree.argtypes = [ctypes.c_int, ctypes.c_int] + [ctypes.c_long] * 3 + [ctypes.c_int] * 2
    for order, bs, dims, nranks in ((1, 1, (5, 4, 6), 1), (1, 3, (3, 3, 7), 3), (2, 1, (3, 2, 5), 2), (3, 1, (2, 2, 4), 4)):
        for rank in range(nranks):
            assert lib.header_modes_agree(order, bs, *dims, rank, nranks) == 0, (order, dims, rank, nranks)


@pytest.mark.parametrize("ptype,dims,nranks,grid,npull", [("elasticity", (9, 8, 10), 1, 7, -1),
                                                          ("elasticity", (6, 7, 12), 2, 9, 2),
                                                          ("poisson", (16, 15, 17), 3, 12, 3)])
def test_balance_plan_covers_every_slice_once_with_equal_work(pt, ptype, dims, nranks, grid, npull):
    """Host plan of the balanced operator split (layout.cpp build_balance_plan): the CTAs' runs tile
    the slice order without gap or overlap, puller runs cover exactly the ghost-reading positions,
    and no run holds more k-steps than its equal share plus one slice."""
    for rank in range(nranks):
        P = pt.host.Problem(ptype, 1, *dims, rank, nranks)
        L = pt.abi.p1_layout(P["dofmap"], P.n_owned, P["rowptr"], P["cols"])
        order, n_int = pt.abi.slice_order(P.n_owned, P["rowptr"], P["cols"])
        S = L["n_slices"]
        use_pull = npull if nranks > 1 else -1
        ou, begin = pt.abi.balance_plan(L["mat_off"], order, n_int, grid, use_pull)
        w = (L["mat_off"][order + 1] - L["mat_off"][order]) // 32
        assert np.array_equal(ou, np.concatenate([[0], np.cumsum(w)]))
        groups = [(n_int, S, use_pull), (0, n_int, grid - use_pull)] if use_pull >= 0 else [(0, S, grid)]
        pos = 0
        for a, b, ctas in groups:
            run = begin[pos:pos + ctas + 1]
            pos += ctas + 1
            assert run[0] == a and run[-1] == b and np.all(np.diff(run) >= 0)
            units = ou[run[1:]] - ou[run[:-1]]
            share = (ou[b] - ou[a]) / ctas
            assert units.max() <= share + w.max() + 1
        assert pos == len(begin)
