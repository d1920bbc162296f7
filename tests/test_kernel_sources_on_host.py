"""The P1 walk kernels' *sources* executed on the host (tests/emu/emu_kernels.cpp): one std::thread
per CUDA thread, a barrier for __syncthreads, a plain array for shared memory. This checks what the
numpy restatements cannot: the kernels' own indexing, shared-memory layouts, register-position
logic and epilogues. Every kernel executed here has also run on the B200 (if the harness and the GPU
disagree the harness is wrong); the harness is where new kernels are debugged before GPU time is
spent on them. CPU only; nothing here is part of the product path.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
# PTB_EMU_FMA=1 builds every harness the way nvcc builds the kernels -- products and sums contracted
# into fused multiply-adds -- into its own directory: the 1e-12 parity bounds must hold there too.
FMA = os.environ.get("PTB_EMU_FMA") == "1"
CXXFLAGS = ["-O2", "-mfma", "-ffp-contract=fast"] if FMA else ["-O1"]
BUILD = "_build_fma" if FMA else "_build"
SRC = os.path.join(HERE, "emu", "emu_kernels.cpp")
OUT = os.path.join(HERE, "emu", BUILD, "libemu.so")
CSRC = os.path.join(os.path.dirname(HERE), "performance-test_b200", "csrc")


@pytest.fixture(scope="module")
def emu():
    deps = [SRC] + [os.path.join(CSRC, f) for f in
                    ("assemble_walk.cu", "assemble_gwalk.cu", "assemble_ring.cu", "geom.cuh", "kernels.h", "ctx.h")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        subprocess.run(["/usr/bin/g++", "-std=c++20", *CXXFLAGS, "-fPIC", "-shared", "-pthread", "-w",
                        "-I", cuda_inc, "-o", OUT, SRC], check=True)
    return C.CDLL(OUT)


OUT_FMA = os.path.join(HERE, "emu", "_build", "libemu_fma.so")  # always the FMA build


@pytest.fixture(scope="module")
def emu_fma():
    """The same harness built the way nvcc builds the kernels: products and sums contracted into
    fused multiply-adds (-mfma -ffp-contract=fast). Needs a host CPU with FMA."""
    if "fma" not in open("/proc/cpuinfo").read().split():
        pytest.skip("host CPU without FMA")
    deps = [SRC] + [os.path.join(CSRC, f) for f in
                    ("assemble_walk.cu", "assemble_gwalk.cu", "assemble_ring.cu", "geom.cuh", "kernels.h", "ctx.h")]
    if not os.path.exists(OUT_FMA) or any(os.path.getmtime(d) > os.path.getmtime(OUT_FMA) for d in deps):
        os.makedirs(os.path.dirname(OUT_FMA), exist_ok=True)
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        subprocess.run(["/usr/bin/g++", "-std=c++20", "-O2", "-mfma", "-ffp-contract=fast", "-fPIC", "-shared",
                        "-pthread", "-w", "-I", cuda_inc, "-o", OUT_FMA, SRC], check=True)
    return C.CDLL(OUT_FMA)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _inputs(pt, P):
    L = pt.abi.p1_layout(P["dofmap"], P.n_owned, P["rowptr"], P["cols"])
    nl = P.n_owned + P.n_ghost
    xdof = np.zeros((nl, 4))
    xdof[:, :3] = P["dof_x"].reshape(-1, 3)
    bc = np.zeros(nl, np.uint8)
    bc[P["bc_dofs"]] = 1
    return L, np.ascontiguousarray(xdof.reshape(-1)), bc


def _sell_to_csr(P, L, vals_sell, bs2):
    rp = P["rowptr"]
    out = np.zeros((rp[-1], bs2))
    for r in range(P.n_owned):
        mo = L["mat_off"][r >> 5]
        for k in range(rp[r + 1] - rp[r]):
            for e in range(bs2):
                out[rp[r] + k, e] = vals_sell[(mo + k * 32) * bs2 + e * 32 + (r & 31)]
    return out.reshape(-1)


def _row_diag(P, ref, bs2):
    rows = np.repeat(np.arange(P.n_owned), np.diff(P["rowptr"]))
    diag = np.zeros(P.n_owned)
    d = P["cols"] == rows
    diag[rows[d]] = np.abs(ref.reshape(-1, bs2)[d]).max(axis=1)
    return np.repeat(diag[rows], bs2)


MATRIX = [(0, "poisson", (5, 4, 6), 0, 1), (1, "poisson", (5, 4, 6), 0, 1), (0, "poisson", (1, 1, 1), 0, 1),
          (0, "poisson", (3, 7, 2), 0, 1), (1, "poisson", (4, 3, 5), 1, 2),
          (6, "poisson", (5, 4, 6), 0, 1), (6, "poisson", (4, 3, 5), 1, 2),
          (2, "elasticity", (4, 3, 3), 0, 1), (2, "elasticity", (1, 1, 2), 0, 1), (2, "elasticity", (3, 3, 4), 1, 2)]


@pytest.mark.parametrize("jitter", [False, True])
@pytest.mark.parametrize("variant,ptype,dims,rank,nranks", MATRIX)
def test_matrix_kernel_sources_reproduce_the_oracle(pt, oracle, emu, perturbed, variant, ptype, dims, rank,
                                                    nranks, jitter):
    P = pt.host.Problem(ptype, 1, *dims, rank, nranks)
    if jitter:
        P = perturbed(P)
    L, xdof, bc = _inputs(pt, P)
    bs2 = P.bs * P.bs
    vals = np.full(int(L["mat_off"][-1]) * bs2, np.nan)
    dinv = np.full(P.n_owned * P.bs, np.nan)
    rp = np.ascontiguousarray(P["rowptr"])
    rc = emu.emu_assemble_matrix(variant, P.n_owned, L["n_slices"], L["max_w"], P.bs, _p(bc), _p(rp),
                                 _p(L["mat_off"]), _p(L["adj_off"]), _p(L["cols"]), _p(xdof),
                                 _p(L["walk"]), _p(L["walk1"]), _p(L["walk1_off"]), _p(vals), _p(dinv))
    assert rc == 0
    assert not np.isnan(vals).any(), "a stored value (padding included) was never written"
    got = _sell_to_csr(P, L, vals, bs2)
    ref = oracle.assemble_matrix(P)
    assert (np.abs(got - ref) / _row_diag(P, ref, bs2)).max() <= 1e-12
    A = ref.reshape(-1, P.bs, P.bs)
    rows = np.repeat(np.arange(P.n_owned), np.diff(P["rowptr"]))
    own = P["cols"] == rows
    d = np.stack([A[own][:, i, i] for i in range(P.bs)], axis=1).reshape(-1)
    assert np.allclose(dinv, 1.0 / d, rtol=1e-12, atol=0)
    # padding entries of the SELL slices must be exact zeros (the SpMV multiplies them)
    mask = np.ones(len(vals), bool)
    for r in range(P.n_owned):
        mo = L["mat_off"][r >> 5]
        for k in range(P["rowptr"][r + 1] - P["rowptr"][r]):
            for e in range(bs2):
                mask[(mo + k * 32) * bs2 + e * 32 + (r & 31)] = False
    assert np.all(vals[mask] == 0.0)
    if variant == 6 and not jitter:
        # the EXACT variant on the lattice: every entry that the oracle (no FMA) computes as an exact
        # zero is an exact zero here too -- the 7-point stencil inside the 15-entry pattern
        assert np.array_equal(got == 0.0, ref == 0.0) and (ref == 0.0).mean() > 0.3


@pytest.mark.parametrize("dims,rank,nranks", [((5, 4, 6), 0, 1), ((7, 3, 5), 1, 2)])
def test_contracted_arithmetic_leaves_residue_where_the_exact_variant_has_zeros(pt, oracle, emu_fma, perturbed, dims,
                                                                                 rank, nranks):
    """What fused multiply-adds do to the lattice operator, shown on the host: with contraction the
    default star-walk kernel leaves rounding residue (~1e-17 of the diagonal) in entries that vanish
    analytically, its EXACT instantiation (cofactor vectors without contraction) gives the oracle's
    exact zeros -- the reason PTB_SPMV_COMPACT=1 selects it. Both stay within 1e-12 of the oracle,
    on the lattice and on a jittered mesh."""
    for jitter in (False, True):
        P = pt.host.Problem("poisson", 1, *dims, rank, nranks)
        if jitter:
            P = perturbed(P)
        L, xdof, bc = _inputs(pt, P)
        ref = oracle.assemble_matrix(P)
        rp = np.ascontiguousarray(P["rowptr"])
        got = {}
        for variant in (0, 6):
            vals = np.full(int(L["mat_off"][-1]), np.nan)
            dinv = np.full(P.n_owned, np.nan)
            assert emu_fma.emu_assemble_matrix(variant, P.n_owned, L["n_slices"], L["max_w"], 1, _p(bc), _p(rp),
                                               _p(L["mat_off"]), _p(L["adj_off"]), _p(L["cols"]), _p(xdof),
                                               _p(L["walk"]), _p(L["walk1"]), _p(L["walk1_off"]), _p(vals),
                                               _p(dinv)) == 0
            got[variant] = _sell_to_csr(P, L, vals, 1)
            assert (np.abs(got[variant] - ref) / _row_diag(P, ref, 1)).max() <= 1e-12
        if not jitter:
            zeros = ref == 0.0
            assert zeros.mean() > 0.3                                   # the 7-point stencil in 15 entries
            for exact, plain in ((6, 0),):                              # star walk
                assert np.array_equal(got[exact] == 0.0, zeros)         # exact variant: exact zeros
                residue = np.abs(got[plain][zeros]) / _row_diag(P, ref, 1)[zeros]
                assert np.count_nonzero(residue) > 0 and residue.max() < 1e-15   # contracted: residue


VECTOR = [("poisson", (5, 4, 6), 0, 1, 4), ("poisson", (1, 1, 1), 0, 1, 1), ("poisson", (4, 3, 5), 1, 2, 4),
          ("elasticity", (3, 4, 3), 0, 1, 4), ("elasticity", (2, 2, 5), 1, 2, 8)]


@pytest.mark.parametrize("jitter", [False, True])
@pytest.mark.parametrize("ptype,dims,rank,nranks,warps", VECTOR)
def test_vector_kernel_source_reproduces_the_oracle(pt, oracle, emu, perturbed, ptype, dims, rank, nranks,
                                                    warps, jitter):
    P = pt.host.Problem(ptype, 1, *dims, rank, nranks)
    if jitter:
        P = perturbed(P)
    L, xdof, bc = _inputs(pt, P)
    b = np.full(P.n_owned * P.bs, np.nan)
    f = np.ascontiguousarray(P["f"])
    rc = emu.emu_assemble_vector(P.bs, warps, P.n_owned, L["n_slices"], L["max_w"], _p(bc),
                                 _p(L["mat_off"]), _p(L["cols"]), _p(xdof), _p(f), _p(L["walk1"]),
                                 _p(L["walk1_off"]), _p(b))
    assert rc == 0 and not np.isnan(b).any()
    b_ref = oracle.assemble_vector(P)
    keep = np.ones(P.n_owned, bool)
    if ptype == "poisson":  # the g v ds facet term belongs to another kernel
        touched = np.zeros(P.n_owned + P.n_ghost, bool)
        dm = P["dofmap"].reshape(-1, 4)
        for c, lf in zip(P["facet_cells"], P["facet_local"]):
            touched[[dm[c, v] for v in range(4) if v != lf]] = True
        keep = ~touched[:P.n_owned]
    keep = np.repeat(keep, P.bs)
    if keep.any():
        assert np.abs(b - b_ref)[keep].max() <= 1e-12 * np.abs(b_ref).max()


# ---- the persistent CG loop (csrc/cg.cu cg_loop): G copies of the harness = G concurrent CTAs ------

CG_SRC = os.path.join(HERE, "emu", "emu_cg.cpp")
CGSTATE = np.dtype([("py", "f8"), ("rr", "f8"), ("rz", "f8"), ("rz_old", "f8"), ("rnorm0", "f8"),
                    ("rtol2", "f8"), ("rnorm", "f8"), ("alpha", "f8"), ("k", "i4"), ("conv", "i4")])


@pytest.fixture(scope="module")
def emucg():
    import shutil
    base = os.path.join(HERE, "emu", BUILD, "libemucg.so")
    deps = [CG_SRC] + [os.path.join(CSRC, f) for f in
                       ("cg.cu", "reduce.cuh", "peer.cuh", "peer.h", "sync_ops.cuh", "kernels.h", "ctx.h")]
    if not os.path.exists(base) or any(os.path.getmtime(d) > os.path.getmtime(base) for d in deps):
        os.makedirs(os.path.dirname(base), exist_ok=True)
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        subprocess.run(["/usr/bin/g++", "-std=c++20", *CXXFLAGS, "-fPIC", "-shared", "-pthread", "-w",
                        "-I", cuda_inc, "-o", base, CG_SRC], check=True)
    libs = []
    for g in range(16):  # one copy per CTA: `__shared__` variables are statics of the copy
        path = base.replace(".so", f"_{os.environ.get('PYTEST_XDIST_WORKER', 'w')}_{g}.so")  # never rewrite a copy another worker has mapped
        shutil.copyfile(base, path)
        libs.append(C.CDLL(path))
    assert libs[0].emu_cgstate_size() == CGSTATE.itemsize
    return libs


def _balance_plan(mat_off, order, groups):
    """The host side of the balanced operator split (cg.cu ensure_balance) restated: unit prefix over
    the slice order and, per (first position, last position, CTAs) group, the CTAs' runs with equal
    units up to one slice. Returns (ounit, begin), begin = None where a run exceeds 64 slices."""
    order = np.asarray(order)
    w = (np.asarray(mat_off)[order + 1] - np.asarray(mat_off)[order]) // 32
    ou = np.concatenate([[0], np.cumsum(w)]).astype(np.int32)
    begin = []
    for a, b, ctas in groups:
        prev = a
        for t in range(ctas + 1):
            target = int(ou[a]) + (int(ou[b]) - int(ou[a])) * t // ctas
            i = a + int(np.searchsorted(ou[a:b + 1], target, side="left"))
            if i > a and target - ou[i - 1] < ou[i] - target:
                i -= 1
            i = max(i, prev)
            if t == ctas:
                i = b
            if t > 0 and i - prev > 64:
                return ou, None
            begin.append(i)
            prev = i
    return ou, np.array(begin, dtype=np.int32)


@pytest.mark.parametrize("balanced", [False, True, "resident"])
@pytest.mark.parametrize("ptype,dims,precond,grid", [("poisson", (6, 5, 7), "jacobi", 3),
                                                     ("poisson", (6, 5, 7), "none", 1),
                                                     ("poisson", (9, 2, 2), "jacobi", 4),
                                                     ("elasticity", (3, 4, 3), "jacobi", 2)])
def test_persistent_cg_loop_source_reproduces_the_oracle(pt, oracle, emucg, ptype, dims, precond, grid,
                                                         balanced):
    """balanced = True runs the instantiation that cuts the CTA's slices into equal k-step ranges per
    warp (spmv_cta_balanced): slices split between warps are summed through shared memory;
    "resident" additionally keeps x and r of the CTA's own rows in shared memory for the whole solve."""
    import threading
    P = pt.host.Problem(ptype, 1, *dims)
    bs, n, rtol = P.bs, P.n_owned * P.bs, 1e-8
    A, b = oracle.assemble_matrix(P), oracle.assemble_vector(P)
    x_ref, k_ref, _ = oracle.cg(bs, P.n_owned, P["rowptr"], P["cols"], A, b, kmax=500, rtol=rtol,
                                precond=precond)
    L = pt.abi.p1_layout(P["dofmap"], P.n_owned, P["rowptr"], P["cols"])
    bs2 = bs * bs
    vals = np.zeros(int(L["mat_off"][-1]) * bs2)
    rp, Ab = P["rowptr"], A.reshape(-1, bs2)
    for r in range(P.n_owned):
        mo = L["mat_off"][r >> 5]
        for k in range(rp[r + 1] - rp[r]):
            vals[(mo + k * 32) * bs2 + np.arange(bs2) * 32 + (r & 31)] = Ab[rp[r] + k]
    cdelta, xoff, colsx = pt.abi.compressed_columns(P.n_owned, P.n_owned + P.n_ghost, rp, P["cols"],
                                                    int(L["mat_off"][-1]))
    order = np.arange(L["n_slices"], dtype=np.int32)
    rows = np.repeat(np.arange(P.n_owned), np.diff(rp))
    own = P["cols"] == rows
    diag = np.stack([A.reshape(-1, bs, bs)[own][:, i, i] for i in range(bs)], axis=1).reshape(-1)
    dinv = 1.0 / diag if precond == "jacobi" else np.ones(n)
    x, r, y = np.zeros(n), b.copy(), np.zeros(n)
    p = dinv * r
    st = np.zeros(2, dtype=CGSTATE)
    rr, rz = float(r @ r), float(r @ (dinv * r))
    st[1] = (0.0, rr, rz, rz, rr, rtol * rtol, rr, 0.0, 0, 0)
    slots = np.zeros(4 * (grid + 1), np.uint64)   # LL arrival records of the CTAs + the release record
    args = [P.n_owned, L["n_slices"], _p(L["mat_off"]), _p(L["cols"]), _p(vals), _p(cdelta), _p(colsx),
            _p(xoff), _p(order), _p(dinv), _p(r), _p(p), _p(x), _p(y), _p(st), _p(slots), 500]
    ou, begin = _balance_plan(L["mat_off"], order, [(0, L["n_slices"], grid)]) if balanced else (None, None)
    assert not balanced or begin is not None
    if balanced:   # the product's host plan (layout.cpp build_balance_plan) is the same plan
        ou_p, begin_p = pt.abi.balance_plan(L["mat_off"], order, L["n_slices"], grid)
        assert np.array_equal(ou, ou_p) and np.array_equal(begin, begin_p)
    res_cap = 0
    if balanced == "resident":
        res_cap = int(np.diff(begin).max()) * 32 * bs
    args += [_p(ou) if balanced else None, _p(begin) if balanced else None, res_cap]
    threads = [threading.Thread(target=emucg[g].emu_cg_loop_block, args=[bs, g, grid] + args)
               for g in range(grid)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=240)
        assert not t.is_alive(), "grid barrier did not release"
    fin = st[int(np.argmax(st["k"]))]
    assert fin["conv"] == 1 and abs(int(fin["k"]) - k_ref) <= 1
    assert np.sqrt(fin["rnorm"] / fin["rnorm0"]) < rtol
    assert np.abs(x - x_ref).max() <= 1e-6 * np.abs(x_ref).max()
    # word 0 of every arrival record carries the epoch of the last barrier: lbase + 3 * iterations - 1
    assert np.all(slots[: 4 * grid: 4] >> np.uint64(32) == np.uint64(3 * int(fin["k"])))


def _rank_view(pt, oracle, P):
    """What the peer-loop harness needs from one rank of the stand-in's z-slab partition."""
    # copies: the arrays of a Problem are views into memory the Problem owns
    return dict(bs=P.bs, n_owned=P.n_owned, n_ghost=P.n_ghost, dofmap=np.array(P["dofmap"]), rowptr=np.array(P["rowptr"]),
                cols=np.array(P["cols"]), A=oracle.assemble_matrix(P), b=oracle.assemble_vector(P),
                nbr=np.array(P["nbr_ranks"], dtype=np.int32),
                send_displ=np.array(P["send_displ"]), local_indices=np.array(P["local_indices"]),
                recv_displ=np.array(P["recv_displ"], dtype=np.int32),
                remote=np.array(P["remote_indices"], dtype=np.int32),
                owned_global=P.global_offset + np.arange(P.n_owned))


def _block_partition(pt, oracle, G, blocks):
    """A general partition built in numpy from the serial problem G (P1): vertices owned by the
    (bx, by, bz) block of the lattice they sit in, every rank holds the cells that touch its vertices
    (ghost-cell layer as in the stand-in) and the Scatterer-style halo lists DOLFINx would hand over:
    nbr_ranks ascending, per neighbour the ghosts it owns in ascending global order (remote_indices =
    their local positions) and the owned dofs it ghosts (local_indices). 2 x 2 x 1 blocks give every
    rank three neighbours, 2 x 2 x 2 seven -- the stand-in's z-slabs never more than two."""
    bs, n = G.bs, G.n_owned
    X = np.array(G["dof_x"]).reshape(-1, 3)
    dims = np.array([G.nx, G.ny, G.nz])
    idx = np.rint(X * dims).astype(np.int64)
    B = np.array(blocks)
    blk = np.minimum(idx * B // (dims + 1), B - 1)
    owner = (blk[:, 2] * B[1] + blk[:, 1]) * B[0] + blk[:, 0]
    dm = np.array(G["dofmap"]).reshape(-1, 4)
    rp, cl = np.array(G["rowptr"]), np.array(G["cols"])
    A, b = oracle.assemble_matrix(G).reshape(-1, bs * bs), oracle.assemble_vector(G).reshape(-1, bs)
    nranks = int(B.prod())
    views = []
    for q in range(nranks):
        owned = np.flatnonzero(owner == q)
        cells = np.flatnonzero((owner[dm] == q).any(axis=1))
        ghosts = np.setdiff1d(np.unique(dm[cells]), owned)
        l2g = np.concatenate([owned, ghosts])
        g2l = np.full(n, -1, np.int64)
        g2l[l2g] = np.arange(len(l2g))
        rowptr, cols, vals = [0], [], []
        for g in owned:
            c = g2l[cl[rp[g]:rp[g + 1]]]
            assert (c >= 0).all()              # every neighbour of an owned vertex is in a local cell
            o = np.argsort(c)
            cols.append(c[o])
            vals.append(A[rp[g]:rp[g + 1]][o])
            rowptr.append(rowptr[-1] + len(c))
        views.append(dict(bs=bs, n_owned=len(owned), n_ghost=len(ghosts), dofmap=g2l[dm[cells]].astype(np.int32).reshape(-1),
                          rowptr=np.array(rowptr, np.int64), cols=np.concatenate(cols).astype(np.int32),
                          A=np.concatenate(vals).reshape(-1), b=b[owned].reshape(-1), owned_global=owned,
                          ghosts=ghosts, g2l=g2l))
    for q, v in enumerate(views):
        gown = owner[v["ghosts"]]
        v["nbr"] = np.unique(gown).astype(np.int32)
        recv, remote, send, local = [0], [], [0], []
        for nb in v["nbr"]:
            mine = np.flatnonzero(gown == nb)                       # ghosts of q owned by nb, ascending global
            remote.append(v["n_owned"] + mine)
            recv.append(recv[-1] + len(mine))
            theirs = views[nb]["ghosts"][owner[views[nb]["ghosts"]] == q]   # what nb ghosts from q
            local.append(v["g2l"][theirs])
            send.append(send[-1] + len(theirs))
        v["recv_displ"], v["remote"] = np.array(recv, np.int32), np.concatenate(remote).astype(np.int32)
        v["send_displ"], v["local_indices"] = np.array(send), np.concatenate(local)
    for q, v in enumerate(views):   # the relation is symmetric and the two sides agree entry by entry
        for j, nb in enumerate(v["nbr"]):
            o = views[nb]
            i = list(o["nbr"]).index(q)
            sent = o["owned_global"][o["local_indices"][o["send_displ"][i]:o["send_displ"][i + 1]]]
            got = np.concatenate([v["owned_global"], v["ghosts"]])[v["remote"][v["recv_displ"][j]:v["recv_displ"][j + 1]]]
            assert np.array_equal(sent, got)
    return views


def _peer_loop(pt, emucg, views, x_ref, k_ref, grid, balanced, rtol):
    """nranks x grid CTAs of cg_loop<BS, true> running at once in one address space (one copy of the
    harness per CTA): CTA 0 of a rank is the puller (publishes 'p is ready', waits for its neighbours,
    pulls the ghost values out of their vectors, takes the ghost-reading slices), the others are workers;
    the dot products go through the LL windows of all ranks."""
    import threading
    nranks = len(views)
    assert len(emucg) >= nranks * grid
    assert emucg[0].emu_peerwindow_size() % 8 == 0
    windows = [np.zeros(emucg[0].emu_peerwindow_size() // 8, np.uint64) for _ in range(nranks)]
    win_ptrs = (C.c_void_p * nranks)(*[w.ctypes.data for w in windows])
    R = []
    for V in views:
        bs, n_owned = V["bs"], V["n_owned"]
        n, nl = n_owned * bs, (n_owned + V["n_ghost"]) * bs
        bs2 = bs * bs
        A, b = V["A"], V["b"]
        L = pt.abi.p1_layout(V["dofmap"], n_owned, V["rowptr"], V["cols"])
        vals = np.zeros(int(L["mat_off"][-1]) * bs2)
        rp, Ab = V["rowptr"], A.reshape(-1, bs2)
        for r in range(n_owned):
            mo = L["mat_off"][r >> 5]
            for k in range(rp[r + 1] - rp[r]):
                vals[(mo + k * 32) * bs2 + np.arange(bs2) * 32 + (r & 31)] = Ab[rp[r] + k]
        cdelta, xoff, colsx = pt.abi.compressed_columns(n_owned, n_owned + V["n_ghost"], rp, V["cols"],
                                                        int(L["mat_off"][-1]))
        order, n_int = pt.abi.slice_order(n_owned, rp, V["cols"])
        rows = np.repeat(np.arange(n_owned), np.diff(rp))
        own = V["cols"] == rows
        diag = np.stack([A.reshape(-1, bs, bs)[own][:, i, i] for i in range(bs)], axis=1).reshape(-1)
        d = dict(V=V, L=L, vals=vals, cdelta=cdelta, xoff=xoff, colsx=colsx, order=order, n_int=n_int,
                 dinv=1.0 / diag, x=np.zeros(nl), r=b.copy(), y=np.zeros(n), p=np.zeros(nl),
                 st=np.zeros(2, dtype=CGSTATE), slots=np.zeros(4 * (grid + 1), np.uint64),
                 ready=np.zeros(256, np.uint64), nbr=V["nbr"], recv_displ=V["recv_displ"], remote=V["remote"])
        d["p"][:n] = d["dinv"] * d["r"]
        R.append(d)
    rr = sum(float(d["r"] @ d["r"]) for d in R)
    rz = sum(float(d["r"] @ (d["dinv"] * d["r"])) for d in R)
    for q, d in enumerate(R):
        d["st"][1] = (0.0, rr, rz, rz, rr, rtol * rtol, rr, 0.0, 0, 0)
        src = []
        for nb in d["nbr"]:
            o = R[nb]["V"]
            j = list(o["nbr"]).index(q)
            src.append(np.array(o["local_indices"][o["send_displ"][j]:o["send_displ"][j + 1]]))
        d["src"] = np.concatenate(src).astype(np.int32)
        assert len(d["src"]) == d["recv_displ"][-1]
        d["peer_p"] = (C.c_void_p * len(d["nbr"]))(*[R[nb]["p"].ctypes.data for nb in d["nbr"]])
    threads = []
    for q, d in enumerate(R):
        V, L = d["V"], d["L"]
        for g in range(grid):
            args = [V["bs"], g, grid, q, nranks, win_ptrs, len(d["nbr"]), _p(d["nbr"]), _p(d["recv_displ"]),
                    d["peer_p"], _p(d["remote"]), _p(d["src"]), d["n_int"], 1, _p(d["ready"]),
                    V["n_owned"], L["n_slices"], _p(L["mat_off"]), _p(L["cols"]), _p(d["vals"]),
                    _p(d["cdelta"]), _p(d["colsx"]), _p(d["xoff"]), _p(d["order"]), _p(d["dinv"]),
                    _p(d["r"]), _p(d["p"]), _p(d["x"]), _p(d["y"]), _p(d["st"]), _p(d["slots"]), 500]
            if balanced:   # one puller (ghost-reading slices), the others workers (interior slices)
                if "ou" not in d:
                    d["ou"], d["begin"] = _balance_plan(L["mat_off"], d["order"],
                                                        [(d["n_int"], L["n_slices"], 1), (0, d["n_int"], grid - 1)])
                    assert d["begin"] is not None
                    ou_p, begin_p = pt.abi.balance_plan(L["mat_off"], d["order"], d["n_int"], grid, 1)
                    assert np.array_equal(d["ou"], ou_p) and np.array_equal(d["begin"], begin_p)
                runs = np.concatenate([np.diff(d["begin"][:2]), np.diff(d["begin"][2:])])
                args += [_p(d["ou"]), _p(d["begin"]), int(runs.max()) * 32 * V["bs"] if balanced == "resident" else 0]
            else:
                args += [None, None, 0]
            threads.append(threading.Thread(target=emucg[q * grid + g].emu_cg_loop_block_peer, args=args))
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
        assert not t.is_alive(), "a CTA is stuck (grid barrier, halo flag or window all-reduce)"
    ks = []
    for d in R:
        fin = d["st"][int(np.argmax(d["st"]["k"]))]
        assert fin["conv"] == 1
        ks.append(int(fin["k"]))
        V = d["V"]
        gidx = (V["owned_global"][:, None] * V["bs"] + np.arange(V["bs"])).reshape(-1)
        err = np.abs(d["x"][:V["n_owned"] * V["bs"]] - x_ref[gidx]).max() / np.abs(x_ref).max()
        assert err <= 1e-6, (err, ks, k_ref)
    assert len(set(ks)) == 1 and abs(ks[0] - k_ref) <= 1


@pytest.mark.parametrize("balanced", [False, True, "resident"])
@pytest.mark.parametrize("ptype,dims", [("poisson", (4, 3, 9)), ("elasticity", (2, 3, 5))])
def test_persistent_cg_loop_peer_branch_on_host(pt, oracle, emucg, ptype, dims, balanced):
    """Two ranks (the stand-in's z-slabs) x two CTAs of cg_loop<BS, true>, see _peer_loop."""
    rtol, grid, nranks = 1e-8, 2, 2
    G = pt.host.Problem(ptype, 1, *dims)
    x_ref, k_ref, _ = oracle.cg(G.bs, G.n_owned, G["rowptr"], G["cols"], oracle.assemble_matrix(G),
                                oracle.assemble_vector(G), kmax=500, rtol=rtol, precond="jacobi")
    views = [_rank_view(pt, oracle, pt.host.Problem(ptype, 1, *dims, q, nranks)) for q in range(nranks)]
    _peer_loop(pt, emucg, views, x_ref, k_ref, grid, balanced, rtol)


@pytest.mark.parametrize("ptype,dims,blocks,balanced", [("poisson", (7, 6, 4), (2, 2, 1), False),
                                                        ("elasticity", (5, 5, 3), (2, 2, 1), True),
                                                        ("poisson", (5, 5, 5), (2, 2, 2), False)])
def test_persistent_cg_loop_with_more_than_two_neighbours_on_host(pt, oracle, emucg, ptype, dims, blocks, balanced):
    """The generic halo lists (ptb_set_halo's arguments) with three and seven neighbours per rank: a
    2 x 2 x 1 / 2 x 2 x 2 block partition built in numpy from the serial problem (the stand-in's
    z-slabs never have more than two). Four or eight ranks x two CTAs of cg_loop<BS, true> pull their
    ghosts from all neighbours, reduce over all ranks' windows and must take the serial oracle's
    iterates: same iteration count on every rank, x to 1e-6."""
    rtol, grid = 1e-8, 2
    G = pt.host.Problem(ptype, 1, *dims)
    x_ref, k_ref, _ = oracle.cg(G.bs, G.n_owned, G["rowptr"], G["cols"], oracle.assemble_matrix(G),
                                oracle.assemble_vector(G), kmax=500, rtol=rtol, precond="jacobi")
    views = _block_partition(pt, oracle, G, blocks)
    assert max(len(v["nbr"]) for v in views) == int(np.prod(blocks)) - 1
    _peer_loop(pt, emucg, views, x_ref, k_ref, grid, balanced, rtol)


@pytest.mark.parametrize("variant,ptype", [(0, "poisson"), (2, "elasticity")])
def test_kernel_sources_assemble_partition_independent_bits(pt, emu, variant, ptype):
    """The walk depends on the mesh topology and the ascending cell order only, so an owned row
    gets bit-identical values on every partition -- checked here on the
    kernels' own arithmetic order: 1 rank against the rows of a 3-rank partition."""
    dims = (3, 2, 7)

    def assemble(rank, nranks):
        P = pt.host.Problem(ptype, 1, *dims, rank, nranks)
        L, xdof, bc = _inputs(pt, P)
        bs2 = P.bs * P.bs
        vals = np.full(int(L["mat_off"][-1]) * bs2, np.nan)
        dinv = np.full(P.n_owned * P.bs, np.nan)
        rp = np.ascontiguousarray(P["rowptr"])
        assert emu.emu_assemble_matrix(variant, P.n_owned, L["n_slices"], L["max_w"], P.bs, _p(bc), _p(rp),
                                       _p(L["mat_off"]), _p(L["adj_off"]), _p(L["cols"]), _p(xdof),
                                       _p(L["walk"]), _p(L["walk1"]), _p(L["walk1_off"]), _p(vals),
                                       _p(dinv)) == 0
        csr = _sell_to_csr(P, L, vals, bs2).reshape(-1, bs2)
        glob = np.concatenate([P.global_offset + np.arange(P.n_owned), P["ghost_global"]])
        out = {}
        for r in range(P.n_owned):
            for k in range(rp[r], rp[r + 1]):
                out[(int(glob[r]), int(glob[P["cols"][k]]))] = csr[k].tobytes()
        return out

    whole = assemble(0, 1)
    parts = {}
    for q in range(3):
        parts.update(assemble(q, 3))
    assert parts.keys() == whole.keys()
    assert all(parts[k] == whole[k] for k in whole)


# ---- elasticity P1 along the edge rings (csrc/assemble_ring.cu) ---------------------------------------

def _ring_assemble(pt, lib, P, warps):
    L, xdof, bc = _inputs(pt, P)
    ring_off, ring_ns, ring = pt.abi.p1_rings(P["dofmap"], P.n_owned, P["rowptr"], P["cols"], int(L["mat_off"][-1]))
    vals = np.full(int(L["mat_off"][-1]) * 9, np.nan)
    dinv = np.full(P.n_owned * 3, np.nan)
    rp = np.ascontiguousarray(P["rowptr"])
    assert lib.emu_assemble_matrix_ring(warps, P.n_owned, L["n_slices"], L["max_w"], _p(bc), _p(rp),
                                        _p(L["mat_off"]), _p(L["cols"]), _p(xdof), _p(ring), _p(ring_off),
                                        _p(ring_ns), int(np.diff(ring_off).max()) // 32, _p(vals), _p(dinv)) == 0
    return L, vals, dinv


@pytest.mark.parametrize("jitter", [False, True])
@pytest.mark.parametrize("dims,rank,nranks,warps", [((4, 3, 3), 0, 1, 4), ((1, 1, 2), 0, 1, 1), ((3, 3, 4), 1, 2, 4),
                                                    ((2, 5, 3), 2, 3, 1), ((9, 2, 2), 0, 1, 4)])
def test_ring_kernel_source_reproduces_the_oracle(pt, oracle, emu, emu_fma, perturbed, dims, rank, nranks, warps,
                                                  jitter):
    """assemble_matrix_p1_ring3 on the host, plain and FMA-contracted build: every stored block within
    1e-12 of its row's diagonal of the oracle's quadrature kernel (the diagonal block comes from
    T_ii = -sum_j T_ij), 1/diag, exact zeros in the SELL padding."""
    P = pt.host.Problem("elasticity", 1, *dims, rank, nranks)
    if jitter:
        P = perturbed(P)
    ref = oracle.assemble_matrix(P)
    for lib in (emu, emu_fma):
        L, vals, dinv = _ring_assemble(pt, lib, P, warps)
        assert not np.isnan(vals).any(), "a stored value (padding included) was never written"
        got = _sell_to_csr(P, L, vals, 9)
        assert (np.abs(got - ref) / _row_diag(P, ref, 9)).max() <= 1e-12
        A = ref.reshape(-1, 3, 3)
        rows = np.repeat(np.arange(P.n_owned), np.diff(P["rowptr"]))
        own = P["cols"] == rows
        d = np.stack([A[own][:, i, i] for i in range(3)], axis=1).reshape(-1)
        assert np.allclose(dinv, 1.0 / d, rtol=1e-12, atol=0)
        mask = np.ones(len(vals), bool)
        for r in range(P.n_owned):
            mo = L["mat_off"][r >> 5]
            for k in range(P["rowptr"][r + 1] - P["rowptr"][r]):
                for e in range(9):
                    mask[(mo + k * 32) * 9 + e * 32 + (r & 31)] = False
        assert np.all(vals[mask] == 0.0)


def test_ring_kernel_source_is_partition_independent_off_the_diagonal(pt, emu):
    """The chains depend on the mesh topology and the ascending cell order only: every off-diagonal
    block of an owned row has the same bits on 1 rank and on a 3-rank partition; the diagonal block
    is summed in local column order (ghost columns last) and agrees to a few ulp."""
    dims = (3, 2, 7)

    def assemble(rank, nranks):
        P = pt.host.Problem("elasticity", 1, *dims, rank, nranks)
        L, vals, _ = _ring_assemble(pt, emu, P, 4)
        csr = _sell_to_csr(P, L, vals, 9).reshape(-1, 9)
        glob = np.concatenate([P.global_offset + np.arange(P.n_owned), P["ghost_global"]])
        rp = P["rowptr"]
        return {(int(glob[r]), int(glob[P["cols"][k]])): csr[k].copy() for r in range(P.n_owned)
                for k in range(rp[r], rp[r + 1])}

    whole = assemble(0, 1)
    parts = {}
    for q in range(3):
        parts.update(assemble(q, 3))
    assert parts.keys() == whole.keys()
    for (r, c), blk in whole.items():
        if r != c:
            assert parts[(r, c)].tobytes() == blk.tobytes()
        else:
            assert np.abs(parts[(r, c)] - blk).max() <= 1e-14 * np.abs(blk).max()


# ---- an unstructured mesh: nothing of the Kuhn box's regularity --------------------------------------

class _Unstructured:
    """Delaunay tetrahedralisation of random points in the unit cube (scipy), vertices in random order:
    stars of 8-40 cells, edge rings of 3-9 cells, fans on the hull, both orientations of the cells,
    mixed ring lengths inside every slice. Slivers are dropped, which also leaves a few edges whose
    cells form more than one fan (chain restarts). Duck-typed like pt.host.Problem for the oracle."""

    def __init__(self, ptype, n_points, seed, order=1):
        from scipy.spatial import Delaunay
        rng = np.random.default_rng(seed)
        corners = np.array([[i, j, k] for i in (0.0, 1.0) for j in (0.0, 1.0) for k in (0.0, 1.0)])
        X = np.concatenate([corners, rng.uniform(0.0, 1.0, size=(n_points, 3))])
        X = X[rng.permutation(len(X))]
        tets = Delaunay(X).simplices.astype(np.int32)
        E = X[tets[:, 1:]] - X[tets[:, :1]]
        vol = np.abs(np.linalg.det(E)) / 6.0
        h = np.linalg.norm(E, axis=2).max(axis=1)
        tets = tets[vol > 0.02 * h ** 3]                     # no slivers
        used = np.unique(tets)
        assert len(used) == len(X)                           # every vertex keeps a cell
        self.problem_type, self.order, self.bs = ptype, order, (3 if ptype == "elasticity" else 1)
        dofmap, DX = tets, X
        if order == 2:   # one dof per edge (midpoint), local dof 4 + e on edge e of the reference tetrahedron
            tet_edges = [(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)]
            edge_id, mids, rowsd = {}, [], []
            for t in tets:
                loc = list(t)
                for a, b in tet_edges:
                    key = (min(t[a], t[b]), max(t[a], t[b]))
                    if key not in edge_id:
                        edge_id[key] = len(X) + len(edge_id)
                        mids.append(0.5 * (X[key[0]] + X[key[1]]))
                    loc.append(edge_id[key])
                rowsd.append(loc)
            dofmap, DX = np.array(rowsd, np.int32), np.concatenate([X, np.array(mids)])
        elif order == 3:   # two dofs per edge (ordered from the lower global vertex), one per face
            tets = np.sort(tets, axis=1)      # ascending vertices: every local edge runs low -> high, as in the Kuhn box
            tet_edges = [(2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1)]
            tet_faces = [(1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2)]
            edge_id, face_id, rowsd = {}, {}, []
            for t in tets:
                loc = list(t)
                for a, b in tet_edges:
                    key = (t[a], t[b])
                    if key not in edge_id:
                        edge_id[key] = len(edge_id)
                    loc += [len(X) + 2 * edge_id[key], len(X) + 2 * edge_id[key] + 1]
                rowsd.append(loc)
            ne = len(edge_id)
            for t, loc in zip(tets, rowsd):
                for fv in tet_faces:
                    key = tuple(t[list(fv)])
                    if key not in face_id:
                        face_id[key] = len(face_id)
                    loc.append(len(X) + 2 * ne + face_id[key])
            # (the points only place the Dirichlet marker here: f is given per dof)
            dofmap = np.array(rowsd, np.int32)
            epts = np.zeros((2 * ne + len(face_id), 3))
            for key, e in edge_id.items():
                epts[2 * e] = (2 * X[key[0]] + X[key[1]]) / 3
                epts[2 * e + 1] = (X[key[0]] + 2 * X[key[1]]) / 3
            for key, f_ in face_id.items():
                epts[2 * ne + f_] = X[list(key)].mean(axis=0)
            DX = np.concatenate([X, epts])
        else:
            assert order == 1
        self.n_cells, self.n_owned, self.n_ghost, self.nd = len(tets), len(DX), 0, dofmap.shape[1]
        rows = [set() for _ in range(len(DX))]
        for t in dofmap:
            for v in t:
                rows[v].update(int(u) for u in t)
        cols = [np.array(sorted(r), np.int32) for r in rows]
        X_geo, X = X, DX
        self._d = {"x": np.ascontiguousarray(X_geo.reshape(-1)), "dof_x": np.ascontiguousarray(DX.reshape(-1)),
                   "x_dofmap": np.ascontiguousarray(tets.reshape(-1)), "dofmap": np.ascontiguousarray(dofmap.reshape(-1)),
                   "rowptr": np.concatenate([[0], np.cumsum([len(c) for c in cols])]).astype(np.int64),
                   "cols": np.concatenate(cols),
                   "bc_dofs": np.flatnonzero(X[:, 0] < 0.15).astype(np.int32),
                   "f": np.ascontiguousarray(rng.standard_normal(len(X) * self.bs)), "g": np.zeros(0),
                   "facet_cells": np.zeros(0, np.int32), "facet_local": np.zeros(0, np.int32)}
        if ptype == "poisson":   # g v ds over the exterior facets: the faces that belong to one cell only
            faces = {}
            for c, t in enumerate(tets):
                for lf in range(4):   # local facet lf is opposite to local vertex lf
                    faces.setdefault(tuple(sorted(int(v) for i, v in enumerate(t) if i != lf)), []).append((c, lf))
            ext = sorted(v[0] for v in faces.values() if len(v) == 1)
            self._d["facet_cells"] = np.array([c for c, _ in ext], np.int32)
            self._d["facet_local"] = np.array([lf for _, lf in ext], np.int32)
            self._d["g"] = np.ascontiguousarray(rng.standard_normal(len(X)))

    def __getitem__(self, k):
        return self._d[k]


@pytest.mark.parametrize("seed,n_points", [(1, 60), (2, 150)])
def test_ring_and_walk_kernel_sources_on_an_unstructured_mesh(pt, oracle, emu, seed, n_points):
    """The elasticity matrix kernels (edge rings and star walk) and their host-built maps on a Delaunay
    mesh: same oracle, same 1e-12 bound, and the two kernels agree with each other."""
    P = _Unstructured("elasticity", n_points, seed)
    ref = oracle.assemble_matrix(P)
    scale = _row_diag(P, ref, 9)
    # chains: for row i and column k the chain's cells are exactly the mesh cells that hold both
    dm = P["dofmap"].reshape(-1, 4)
    rp, cols = P["rowptr"], P["cols"]
    L = pt.abi.p1_layout(P["dofmap"], P.n_owned, rp, cols)
    ring_off, ring_ns, ring = pt.abi.p1_rings(P["dofmap"], P.n_owned, rp, cols, int(L["mat_off"][-1]))
    lens, restarts = [], 0
    for r in range(P.n_owned):
        s, lane = r >> 5, r & 31
        k0, base = int(L["mat_off"][s]) // 32, int(ring_off[s]) + lane
        row_cols = cols[rp[r]:rp[r + 1]]
        mine = dm[(dm == r).any(axis=1)]
        for k in range(int(L["mat_off"][s + 1] - L["mat_off"][s]) // 32):
            ns = int(ring_ns[k0 + k])
            by = [(int(ring[base + (t // 4) * 32]) >> (8 * (t % 4))) & 0xFF for t in range(ns)]
            base += ((ns + 3) // 4) * 32
            if k >= len(row_cols) or row_cols[k] == r:
                assert all(b == 0x80 for b in by)
                continue
            while by and by[-1] == 0x80:
                by.pop()
            j = row_cols[k]
            want = sorted(tuple(sorted(int(v) for v in c if v != r and v != j)) for c in mine if j in c)
            got = [tuple(sorted((int(row_cols[by[t - 1] & 0x7F]), int(row_cols[by[t] & 0x7F]))))
                   for t in range(1, len(by)) if not by[t] & 0x80]
            assert sorted(got) == want
            lens.append(len(want))
            restarts += sum(1 for t in range(1, len(by)) if by[t] & 0x80)
    assert min(lens) <= 2 and max(lens) >= 7            # nothing like the 4- and 6-rings of the Kuhn box
    for warps in (1, 4):
        Lr, vals, dinv = _ring_assemble(pt, emu, P, warps)
        assert not np.isnan(vals).any()
        ring_csr = _sell_to_csr(P, Lr, vals, 9)
        assert (np.abs(ring_csr - ref) / scale).max() <= 1e-12
    _, xdof, bc = _inputs(pt, P)
    vals = np.full(int(L["mat_off"][-1]) * 9, np.nan)
    dinv3 = np.full(P.n_owned * 3, np.nan)
    rp64 = np.ascontiguousarray(rp)
    assert emu.emu_assemble_matrix(2, P.n_owned, L["n_slices"], L["max_w"], 3, _p(bc), _p(rp64), _p(L["mat_off"]),
                                   _p(L["adj_off"]), _p(L["cols"]), _p(xdof), _p(L["walk"]), _p(L["walk1"]),
                                   _p(L["walk1_off"]), _p(vals), _p(dinv3)) == 0
    walk_csr = _sell_to_csr(P, L, vals, 9)
    assert (np.abs(walk_csr - ref) / scale).max() <= 1e-12
    assert (np.abs(walk_csr - ring_csr) / scale).max() <= 1e-13
    assert np.allclose(dinv, dinv3, rtol=1e-12, atol=0)


@pytest.mark.parametrize("ptype,seed,n_points", [("poisson", 1, 60), ("poisson", 2, 150), ("elasticity", 3, 100)])
def test_scalar_walk_and_vector_kernel_sources_on_an_unstructured_mesh(pt, oracle, emu, emup1, ptype, seed, n_points):
    """The scalar star-walk matrix kernel (rows of at most 32 columns) and the P1 cell-vector kernel
    (one thread per block row) on the Delaunay mesh, against the oracle."""
    P = _Unstructured(ptype, n_points, seed)
    L, xdof, bc = _inputs(pt, P)
    assert L["max_w"] <= 32
    rp64 = np.ascontiguousarray(P["rowptr"])
    if ptype == "poisson":
        ref = oracle.assemble_matrix(P)
        vals, dinv = np.full(int(L["mat_off"][-1]), np.nan), np.full(P.n_owned, np.nan)
        assert emu.emu_assemble_matrix(0, P.n_owned, L["n_slices"], L["max_w"], 1, _p(bc), _p(rp64), _p(L["mat_off"]),
                                       _p(L["adj_off"]), _p(L["cols"]), _p(xdof), _p(L["walk"]), _p(L["walk1"]),
                                       _p(L["walk1_off"]), _p(vals), _p(dinv)) == 0
        assert (np.abs(_sell_to_csr(P, L, vals, 1) - ref) / _row_diag(P, ref, 1)).max() <= 1e-12
    b = np.full(P.n_owned * P.bs, np.nan)
    f = np.ascontiguousarray(P["f"])
    assert emu.emu_assemble_vector(P.bs, 4, P.n_owned, L["n_slices"], L["max_w"], _p(bc), _p(L["mat_off"]),
                                   _p(L["cols"]), _p(xdof), _p(f), _p(L["walk1"]), _p(L["walk1_off"]), _p(b)) == 0
    assert not np.isnan(b).any()
    if ptype == "poisson":   # + g v ds over the hull (assemble_facets_p1 through the facet-row lists)
        assert len(P["facet_cells"]) > 20
        ids, ptr, ent = pt.abi.facet_rows(P["facet_cells"], P["facet_local"], P["dofmap"], 4, 1, P.n_owned)
        xyz4 = np.zeros((P.n_owned, 4))
        xyz4[:, :3] = P["x"].reshape(-1, 3)
        xyz4 = np.ascontiguousarray(xyz4.reshape(-1))
        xd, dm, g = np.ascontiguousarray(P["x_dofmap"]), np.ascontiguousarray(P["dofmap"]), np.ascontiguousarray(P["g"])
        assert emup1.emu_p1_facets(len(ids), _p(xyz4), _p(xd), _p(dm), _p(bc), _p(ids), _p(ptr), _p(ent), _p(g), _p(b)) == 0
    b_ref = oracle.assemble_vector(P)
    assert np.abs(b - b_ref).max() <= 1e-12 * np.abs(b_ref).max()


@pytest.mark.parametrize("order", [2, 3])
@pytest.mark.parametrize("binned", [0, 1, 2])
def test_p2_p3_matrix_and_vector_kernel_sources_on_an_unstructured_mesh(pt, oracle, emupk, binned, order):
    """The P2 / P3 kernels (csrc/assemble_pk.cu: all slices, row-length classes, per-cell geometry pre-pass;
    cell vector) and their host-built maps on the Delaunay mesh: one dof per edge (P2); two per edge,
    ordered from the lower global vertex, and one per face (P3)."""
    P = _Unstructured("poisson", 60 if order == 2 else 40, 5 if order == 2 else 3, order=order)
    L = pt.abi.pk_layout(P["dofmap"], P.nd, P.n_owned, P["rowptr"], P["cols"])
    nv = len(P["x"]) // 3
    xyz4 = np.zeros((nv, 4))
    xyz4[:, :3] = P["x"].reshape(-1, 3)
    xyz4 = np.ascontiguousarray(xyz4.reshape(-1))
    bc = np.zeros(P.n_owned, np.uint8)
    bc[P["bc_dofs"]] = 1
    vals, dinv = np.full(int(L["mat_off"][-1]), np.nan), np.full(P.n_owned, np.nan)
    rp, xd, dm = np.ascontiguousarray(P["rowptr"]), np.ascontiguousarray(P["x_dofmap"]), np.ascontiguousarray(P["dofmap"])
    rc = emupk.emu_assemble_matrix_pk(binned, P.nd, L["so_bits"], P.n_owned, L["n_slices"], L["max_w"], _p(xyz4), _p(xd),
                                      _p(bc), _p(rp), _p(L["mat_off"]), _p(L["adj_off"]), _p(L["cols"]), _p(L["adj"]),
                                      _p(L["adjso"]), L["n_bins"], _p(L["bin_off"]), _p(L["bin_w"]),
                                      _p(L["bin_slices"]), _p(vals), _p(dinv), C.c_int64(P.n_cells))
    assert rc == 0 and not np.isnan(vals).any() and not np.isnan(dinv).any()
    ref = oracle.assemble_matrix(P)
    assert (np.abs(_sell_to_csr(P, L, vals, 1) - ref) / _row_diag(P, ref, 1)).max() <= 1e-12
    if binned == 0:   # cell vector + g v ds over the hull (assemble_vector_pk, assemble_facets_pk)
        b = np.full(P.n_owned, np.nan)
        f, g = np.ascontiguousarray(P["f"]), np.ascontiguousarray(P["g"])
        ids, ptr, ent = pt.abi.facet_rows(P["facet_cells"], P["facet_local"], P["dofmap"], P.nd, order, P.n_owned)
        assert len(ids) > 20
        assert emupk.emu_assemble_vector_pk(P.nd, P.n_owned, L["n_slices"], _p(xyz4), _p(xd), _p(dm), _p(bc),
                                            _p(L["adj_off"]), _p(L["adj"]), _p(f), len(ids), _p(ids), _p(ptr), _p(ent),
                                            _p(g), _p(b)) == 0
        b_ref = oracle.assemble_vector(P)
        assert not np.isnan(b).any() and np.abs(b - b_ref).max() <= 1e-12 * np.abs(b_ref).max()


@pytest.mark.parametrize("seed,n_points", [(1, 60), (2, 150)])
def test_device_ring_builder_source_on_an_unstructured_mesh(pt, emusu, seed, n_points):
    """setup.cu setup_rings against layout.cpp build_rings on the Delaunay mesh (fans, restarts, rings of
    3-9 cells): the same words, unless a star exceeds the device kernel's 64 cells -- then flag 1."""
    P = _Unstructured("elasticity", n_points, seed)
    L = pt.abi.p1_layout(P["dofmap"], P.n_owned, P["rowptr"], P["cols"])
    ring_off, ring_ns, ring = pt.abi.p1_rings(P["dofmap"], P.n_owned, P["rowptr"], P["cols"], int(L["mat_off"][-1]))
    S, cap = L["n_slices"], int(ring_off[-1])
    dm, rp = np.ascontiguousarray(P["dofmap"], np.int32), np.ascontiguousarray(P["rowptr"], np.int64)
    d_off, d_ns = np.full(S + 1, -1, np.int64), np.full(len(ring_ns), 0xEE, np.uint8)
    d_ring, flags = np.full(max(cap, 1), 0xDEADBEEF, np.uint32), np.full(2, -1, np.int32)
    rc = emusu.emu_setup_p1_rings(C.c_int64(len(dm) // 4), _p(dm), P.n_owned, S, _p(rp), _p(L["mat_off"]),
                                  _p(L["cols"]), 1, C.c_int64(cap), _p(d_off), _p(d_ns), _p(d_ring), _p(flags))
    star = np.bincount(dm, minlength=P.n_owned).max()
    if star > 64:
        assert flags[1] == 1
        return
    assert rc == 0 and flags.tolist() == [0, 0]
    assert np.array_equal(d_off, ring_off) and np.array_equal(d_ns, ring_ns)
    assert np.array_equal(d_ring[:cap], ring[:cap])


# ---- P2 / P3 matrix kernels (csrc/assemble_pk.cu): all slices at once, and bin after bin -------------

PK_SRC = os.path.join(HERE, "emu", "emu_pk.cpp")


@pytest.fixture(scope="module")
def emupk():
    out = os.path.join(HERE, "emu", BUILD, "libemupk.so")
    deps = [PK_SRC] + [os.path.join(CSRC, f) for f in ("assemble_pk.cu", "element_tables.h", "kernels.h", "ctx.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        # static shared arrays must be one per CTA, not one per host thread (see emu_pk.cpp)
        text = open(os.path.join(CSRC, "assemble_pk.cu")).read()
        text = text.replace("extern __shared__", "extern").replace("__shared__", "static")
        copy = os.path.join(os.path.dirname(out), "assemble_pk_emu.cu")
        with open(copy, "w") as f:
            f.write(text)
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        subprocess.run(["/usr/bin/g++", "-std=c++20", *CXXFLAGS, "-fPIC", "-shared", "-pthread", "-w",
                        "-I", cuda_inc, "-I", CSRC, f'-DPTB_EMU_PK_SOURCE="{copy}"', "-o", out, PK_SRC],
                       check=True)
    return C.CDLL(out)


@pytest.mark.parametrize("binned", [0, 1, 2])   # 2 = geometry factors once per cell (cell_geometry_pk) + binned
@pytest.mark.parametrize("order,dims,rank,nranks", [(2, (3, 2, 4), 0, 1), (3, (2, 3, 2), 0, 1),
                                                    (3, (2, 2, 4), 1, 2)])
def test_p2_p3_matrix_kernel_sources_reproduce_the_oracle(pt, oracle, emupk, perturbed, order, dims, rank,
                                                          nranks, binned):
    P = pt.host.Problem("poisson", order, *dims, rank, nranks)
    if binned:  # the binned runs also take the jittered mesh
        P = perturbed(P)
    L = pt.abi.pk_layout(P["dofmap"], P.nd, P.n_owned, P["rowptr"], P["cols"])
    nv = len(P["x"]) // 3
    xyz4 = np.zeros((nv, 4))
    xyz4[:, :3] = P["x"].reshape(-1, 3)
    xyz4 = np.ascontiguousarray(xyz4.reshape(-1))
    bc = np.zeros(P.n_owned + P.n_ghost, np.uint8)
    bc[P["bc_dofs"]] = 1
    vals = np.full(int(L["mat_off"][-1]), np.nan)
    dinv = np.full(P.n_owned, np.nan)
    rp, xd = np.ascontiguousarray(P["rowptr"]), np.ascontiguousarray(P["x_dofmap"])
    if binned:  # the bins partition the slices, every bin is wide enough for its rows
        assert sorted(L["bin_slices"].tolist()) == list(range(L["n_slices"]))
        for b in range(L["n_bins"]):
            for s in L["bin_slices"][L["bin_off"][b]:L["bin_off"][b + 1]]:
                assert (L["mat_off"][s + 1] - L["mat_off"][s]) // 32 <= L["bin_w"][b]
    rc = emupk.emu_assemble_matrix_pk(binned, P.nd, L["so_bits"], P.n_owned, L["n_slices"], L["max_w"],
                                      _p(xyz4), _p(xd), _p(bc), _p(rp), _p(L["mat_off"]), _p(L["adj_off"]),
                                      _p(L["cols"]), _p(L["adj"]), _p(L["adjso"]), L["n_bins"],
                                      _p(L["bin_off"]), _p(L["bin_w"]), _p(L["bin_slices"]), _p(vals), _p(dinv),
                                      C.c_int64(len(xd) // 4))
    assert rc == 0 and not np.isnan(vals).any() and not np.isnan(dinv).any()
    got = _sell_to_csr(P, L, vals, 1)
    ref = oracle.assemble_matrix(P)
    assert (np.abs(got - ref) / _row_diag(P, ref, 1)).max() <= 1e-12


# ---- the ascending-cell-order P1 kernels (csrc/assemble.cu): default for elasticity and for b --------

P1_SRC = os.path.join(HERE, "emu", "emu_p1.cpp")


@pytest.fixture(scope="module")
def emup1():
    out = os.path.join(HERE, "emu", BUILD, "libemup1.so")
    deps = [P1_SRC] + [os.path.join(CSRC, f) for f in ("assemble.cu", "geom.cuh", "kernels.h", "ctx.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        subprocess.run(["/usr/bin/g++", "-std=c++20", *CXXFLAGS, "-fPIC", "-shared", "-pthread", "-w",
                        "-I", cuda_inc, "-o", out, P1_SRC], check=True)
    return C.CDLL(out)


@pytest.mark.parametrize("jitter", [False, True])
@pytest.mark.parametrize("ptype,dims,rank,nranks", [("poisson", (5, 4, 6), 0, 1), ("elasticity", (4, 3, 3), 0, 1),
                                                    ("elasticity", (1, 1, 2), 0, 1), ("elasticity", (3, 3, 4), 1, 2)])
def test_cell_order_kernel_sources_reproduce_the_oracle(pt, oracle, emup1, perturbed, ptype, dims, rank, nranks,
                                                        jitter):
    P = pt.host.Problem(ptype, 1, *dims, rank, nranks)
    if jitter:
        P = perturbed(P)
    L, xdof, bc = _inputs(pt, P)
    bs2 = P.bs * P.bs
    vals = np.full(int(L["mat_off"][-1]) * bs2, np.nan)
    dinv = np.full(P.n_owned * P.bs, np.nan)
    b = np.full(P.n_owned * P.bs, np.nan)
    rp, f = np.ascontiguousarray(P["rowptr"]), np.ascontiguousarray(P["f"])
    assert emup1.emu_p1_matrix(P.bs, P.n_owned, L["n_slices"], L["max_w"], _p(bc), _p(rp), _p(L["mat_off"]),
                               _p(L["adj_off"]), _p(L["cols"]), _p(L["adjrot"]), _p(xdof), _p(vals),
                               _p(dinv)) == 0
    assert emup1.emu_p1_vector(P.bs, P.n_owned, L["n_slices"], L["max_w"], _p(bc), _p(L["mat_off"]),
                               _p(L["adj_off"]), _p(L["cols"]), _p(L["adjrot"]), _p(xdof), _p(f), _p(b)) == 0
    assert not np.isnan(vals).any() and not np.isnan(b).any()
    ref = oracle.assemble_matrix(P)
    assert (np.abs(_sell_to_csr(P, L, vals, bs2) - ref) / _row_diag(P, ref, bs2)).max() <= 1e-12
    if ptype == "poisson":  # + g v ds over the exterior facets (assemble_facets_p1)
        ids, ptr, ent = pt.abi.facet_rows(P["facet_cells"], P["facet_local"], P["dofmap"], 4, 1, P.n_owned)
        nv = len(P["x"]) // 3
        xyz4 = np.zeros((nv, 4))
        xyz4[:, :3] = np.array(P["x"]).reshape(-1, 3)
        xyz4 = np.ascontiguousarray(xyz4.reshape(-1))
        xd, dm, g = (np.ascontiguousarray(P["x_dofmap"]), np.ascontiguousarray(P["dofmap"]),
                     np.ascontiguousarray(P["g"]))
        assert emup1.emu_p1_facets(len(ids), _p(xyz4), _p(xd), _p(dm), _p(bc), _p(ids), _p(ptr), _p(ent),
                                   _p(g), _p(b)) == 0
    b_ref = oracle.assemble_vector(P)
    assert np.abs(b - b_ref).max() <= 1e-12 * np.abs(b_ref).max()


@pytest.mark.parametrize("jitter", [False, True])
@pytest.mark.parametrize("order,dims,rank,nranks", [(2, (3, 2, 4), 0, 1), (3, (2, 3, 2), 0, 1), (2, (2, 2, 5), 1, 2)])
def test_p2_p3_vector_and_facet_kernel_sources_reproduce_the_oracle(pt, oracle, emupk, perturbed, order, dims,
                                                                     rank, nranks, jitter):
    P = pt.host.Problem("poisson", order, *dims, rank, nranks)
    if jitter:
        P = perturbed(P)
    L = pt.abi.pk_layout(P["dofmap"], P.nd, P.n_owned, P["rowptr"], P["cols"])
    ids, ptr, ent = pt.abi.facet_rows(P["facet_cells"], P["facet_local"], P["dofmap"], P.nd, order, P.n_owned)
    nv = len(P["x"]) // 3
    xyz4 = np.zeros((nv, 4))
    xyz4[:, :3] = np.array(P["x"]).reshape(-1, 3)
    xyz4 = np.ascontiguousarray(xyz4.reshape(-1))
    bc = np.zeros(P.n_owned + P.n_ghost, np.uint8)
    bc[P["bc_dofs"]] = 1
    b = np.full(P.n_owned, np.nan)
    xd, dm = np.ascontiguousarray(P["x_dofmap"]), np.ascontiguousarray(P["dofmap"])
    f, g = np.ascontiguousarray(P["f"]), np.ascontiguousarray(P["g"])
    rc = emupk.emu_assemble_vector_pk(P.nd, P.n_owned, L["n_slices"], _p(xyz4), _p(xd), _p(dm), _p(bc),
                                      _p(L["adj_off"]), _p(L["adj"]), _p(f), len(ids), _p(ids), _p(ptr),
                                      _p(ent), _p(g), _p(b))
    assert rc == 0 and not np.isnan(b).any()
    b_ref = oracle.assemble_vector(P)
    assert np.abs(b - b_ref).max() <= 1e-12 * np.abs(b_ref).max()


# ---- zero-column compaction of the scalar operator (csrc/compact.cu) --------------------------------

CP_SRC = os.path.join(HERE, "emu", "emu_compact.cpp")


@pytest.fixture(scope="module")
def emucp():
    out = os.path.join(HERE, "emu", BUILD, "libemucompact.so")
    deps = [CP_SRC] + [os.path.join(CSRC, f) for f in ("compact.cu", "kernels.h", "ctx.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        text = open(os.path.join(CSRC, "compact.cu")).read().replace("__shared__", "static")
        copy = os.path.join(os.path.dirname(out), "compact_emu.cu")
        with open(copy, "w") as f:
            f.write(text)
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        subprocess.run(["/usr/bin/g++", "-std=c++20", *CXXFLAGS, "-fPIC", "-shared", "-pthread", "-w",
                        "-I", cuda_inc, "-I", CSRC, f'-DPTB_EMU_COMPACT_SOURCE="{copy}"', "-o", out, CP_SRC],
                       check=True)
    return C.CDLL(out)


def _sell_spmv(n_rows, mat_off, vals, cdelta, colsx, xoff, p):
    """numpy restatement of spmv_slice<1> on the compressed SELL-32 arrays."""
    y = np.zeros(n_rows)
    for s in range(len(mat_off) - 1):
        mo, w = int(mat_off[s]), int(mat_off[s + 1] - mat_off[s]) // 32
        jx = 0
        for k in range(w):
            d = int(cdelta[mo // 32 + k])
            for lane in range(32):
                r = 32 * s + lane
                if r >= n_rows:
                    continue
                c = int(colsx[int(xoff[s]) + jx * 32 + lane]) if d == -2 ** 31 else r + d
                y[r] += vals[mo + k * 32 + lane] * p[c]
            jx += d == -2 ** 31
    return y


@pytest.mark.parametrize("dims,rank,nranks,jitter,tol", [((6, 5, 7), 0, 1, False, 0.0), ((9, 3, 4), 1, 2, False, 0.0),
                                                         ((5, 4, 6), 0, 1, True, 0.0), ((6, 5, 7), 0, 1, False, 1e-14)])
def test_operator_compaction_source_keeps_the_operator(pt, oracle, emucp, perturbed, dims, rank, nranks, jitter,
                                                       tol):
    P = pt.host.Problem("poisson", 1, *dims, rank, nranks)
    if jitter:
        P = perturbed(P)
    A = oracle.assemble_matrix(P)
    L = pt.abi.p1_layout(P["dofmap"], P.n_owned, P["rowptr"], P["cols"])
    S, rp = L["n_slices"], P["rowptr"]
    rows = np.repeat(np.arange(P.n_owned), np.diff(rp))
    diag = np.zeros(P.n_owned)
    diag[rows[P["cols"] == rows]] = A[P["cols"] == rows]
    if tol > 0:  # what fused multiply-adds leave behind on the GPU: residue instead of exact zeros
        rng = np.random.default_rng(4)
        A = np.where(A == 0.0, 1e-17 * diag[rows] * rng.standard_normal(len(A)), A)
    dinv = 1.0 / diag
    vals = np.zeros(int(L["mat_off"][-1]))
    for r in range(P.n_owned):
        mo = L["mat_off"][r >> 5]
        vals[mo + np.arange(rp[r + 1] - rp[r]) * 32 + (r & 31)] = A[rp[r]:rp[r + 1]]
    cdelta, xoff, colsx = pt.abi.compressed_columns(P.n_owned, P.n_owned + P.n_ghost, rp, P["cols"],
                                                    int(L["mat_off"][-1]))
    cw, cx = np.zeros(S, np.int64), np.zeros(S, np.int64)
    moz, xoz = np.full(S + 1, -1, np.int64), np.full(S + 1, -1, np.int64)
    assert emucp.emu_compact_offsets(P.n_owned, S, _p(L["mat_off"]), _p(vals), _p(cdelta), _p(dinv),
                                     C.c_double(tol), _p(cw), _p(cx), _p(moz), _p(xoz)) == 0
    assert moz[0] == 0 and np.array_equal(np.diff(moz), cw) and np.array_equal(np.diff(xoz), cx)
    vz = np.full(int(moz[-1]), np.nan)
    cdz = np.zeros(int(moz[-1]) // 32, np.int32)
    cxz = np.zeros(max(int(xoz[-1]), 1), np.int32)
    assert emucp.emu_compact_copy(P.n_owned, S, _p(L["mat_off"]), _p(vals), _p(cdelta), _p(dinv),
                                  C.c_double(tol), _p(colsx), _p(xoff), _p(moz), _p(xoz), _p(vz), _p(cdz),
                                  _p(cxz)) == 0
    assert not np.isnan(vz).any()
    p = np.random.default_rng(2).standard_normal(P.n_owned + P.n_ghost)
    y0 = _sell_spmv(P.n_owned, L["mat_off"], vals, cdelta, colsx, xoff, p)
    y1 = _sell_spmv(P.n_owned, moz, vz, cdz, cxz, xoz, p)
    if tol == 0:
        assert np.array_equal(y0, y1)                  # the dropped terms were exact zeros
    else:
        assert np.abs(y0 - y1).max() <= 1e-15 * np.abs(y0).max()
    assert np.abs(y0 - oracle.spmv(1, P.n_owned, rp, P["cols"], A, p)).max() <= 1e-13 * np.abs(y0).max()
    kept = moz[-1] / L["mat_off"][-1]
    if jitter:
        assert kept > 0.9                              # a general mesh has (almost) no exact zeros
    else:
        assert kept < 0.75                             # the lattice operator is the 7-point stencil


# ---- device-side construction of the P1 assembly maps (csrc/setup.cu) -------------------------------

SU_SRC = os.path.join(HERE, "emu", "emu_setup.cpp")


@pytest.fixture(scope="module")
def emusu():
    out = os.path.join(HERE, "emu", BUILD, "libemusetup.so")
    deps = [SU_SRC] + [os.path.join(CSRC, f) for f in ("setup.cu", "kernels.h", "ctx.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        text = open(os.path.join(CSRC, "setup.cu")).read().replace("__shared__", "static")
        copy = os.path.join(os.path.dirname(out), "setup_emu.cu")
        with open(copy, "w") as f:
            f.write(text)
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        subprocess.run(["/usr/bin/g++", "-std=c++20", *CXXFLAGS, "-fPIC", "-shared", "-pthread", "-w",
                        "-I", cuda_inc, "-I", CSRC, f'-DPTB_EMU_SETUP_SOURCE="{copy}"', "-o", out, SU_SRC],
                       check=True)
    return C.CDLL(out)


@pytest.mark.parametrize("shuffle", [0, 1])
@pytest.mark.parametrize("ptype,dims,rank,nranks", [("poisson", (5, 4, 6), 0, 1), ("poisson", (1, 1, 1), 0, 1),
                                                    ("poisson", (33, 2, 1), 0, 1), ("poisson", (4, 3, 5), 1, 2),
                                                    ("poisson", (3, 3, 7), 2, 3), ("elasticity", (3, 4, 3), 0, 1),
                                                    ("elasticity", (2, 2, 5), 1, 2),
                                                    ("poisson", (12, 11, 13), 0, 1)])  # > 1024 rows: scan chunks > 1
def test_device_setup_source_builds_the_host_maps_bit_for_bit(pt, emusu, ptype, dims, rank, nranks, shuffle):
    """adj_off, the rotated slot words and the star walk built by the setup kernels equal the host
    build (common/intmaps.cpp + layout.cpp) word for word, whatever order the fill atomics land in."""
    P = pt.host.Problem(ptype, 1, *dims, rank, nranks)
    L = pt.abi.p1_layout(P["dofmap"], P.n_owned, P["rowptr"], P["cols"])
    S, cap = L["n_slices"], int(L["adj_off"][-1])
    dm = np.ascontiguousarray(P["dofmap"], np.int32)
    rp = np.ascontiguousarray(P["rowptr"], np.int64)
    adj_off = np.full(S + 1, -1, np.int64)
    adjrot = np.full(cap, 0xDEADBEEF, np.uint32)
    walk = np.full(cap, 0xDEADBEEF, np.uint32)
    flags = np.full(2, -1, np.int32)
    rc = emusu.emu_setup_p1(C.c_int64(len(dm) // 4), _p(dm), P.n_owned, S, _p(rp), _p(L["mat_off"]),
                            _p(L["cols"]), shuffle, C.c_int64(cap), _p(adj_off), _p(adjrot), _p(walk), _p(flags))
    assert rc == 0
    assert np.array_equal(adj_off, L["adj_off"])
    assert flags.tolist() == [0, 0]
    assert np.array_equal(adjrot, L["adjrot"])
    assert np.array_equal(walk, L["walk"])


@pytest.mark.parametrize("shuffle", [0, 1])
@pytest.mark.parametrize("dims,rank,nranks", [((3, 4, 3), 0, 1), ((1, 1, 1), 0, 1), ((2, 2, 5), 1, 2), ((33, 2, 1), 0, 1),
                                              ((3, 3, 7), 2, 3), ((12, 11, 13), 0, 1)])
def test_device_setup_source_builds_the_host_edge_rings_bit_for_bit(pt, emusu, dims, rank, nranks, shuffle):
    """ring_off, ring_ns and the ring words of the column-major elasticity kernel built by the setup
    kernels (setup.cu setup_rings<0/1>, setup_ring_words) equal layout.cpp build_rings word for word."""
    P = pt.host.Problem("elasticity", 1, *dims, rank, nranks)
    L = pt.abi.p1_layout(P["dofmap"], P.n_owned, P["rowptr"], P["cols"])
    ring_off, ring_ns, ring = pt.abi.p1_rings(P["dofmap"], P.n_owned, P["rowptr"], P["cols"], int(L["mat_off"][-1]))
    S, cap = L["n_slices"], int(ring_off[-1])
    dm = np.ascontiguousarray(P["dofmap"], np.int32)
    rp = np.ascontiguousarray(P["rowptr"], np.int64)
    d_off = np.full(S + 1, -1, np.int64)
    d_ns = np.full(len(ring_ns), 0xEE, np.uint8)
    d_ring = np.full(max(cap, 1), 0xDEADBEEF, np.uint32)
    flags = np.full(2, -1, np.int32)
    rc = emusu.emu_setup_p1_rings(C.c_int64(len(dm) // 4), _p(dm), P.n_owned, S, _p(rp), _p(L["mat_off"]),
                                  _p(L["cols"]), shuffle, C.c_int64(cap), _p(d_off), _p(d_ns), _p(d_ring), _p(flags))
    assert rc == 0 and flags.tolist() == [0, 0]
    assert np.array_equal(d_off, ring_off) and np.array_equal(d_ns, ring_ns)
    assert np.array_equal(d_ring[:cap], ring[:cap])
    # and the rotated slot words + star walk of setup_adjrot / setup_walk
    capw = int(L["adj_off"][-1])
    adj_off, adjrot, walk = np.full(S + 1, -1, np.int64), np.full(capw, 0xDEADBEEF, np.uint32), np.full(capw, 0xDEADBEEF, np.uint32)
    flags[:] = -1
    assert emusu.emu_setup_p1(C.c_int64(len(dm) // 4), _p(dm), P.n_owned, S, _p(rp), _p(L["mat_off"]), _p(L["cols"]), 1,
                              C.c_int64(capw), _p(adj_off), _p(adjrot), _p(walk), _p(flags)) == 0
    assert flags.tolist() == [0, 0] and np.array_equal(adj_off, L["adj_off"])
    assert np.array_equal(adjrot, L["adjrot"]) and np.array_equal(walk, L["walk"])


def test_device_setup_source_flags_a_pattern_that_misses_a_cell_pair(pt, emusu):
    """A (row, column) pair of a cell that is absent from the caller's pattern must raise flag 0
    (the host build refuses the pattern: ptb_set_pattern 'pair is missing')."""
    P = pt.host.Problem("poisson", 1, 3, 3, 3)
    L = pt.abi.p1_layout(P["dofmap"], P.n_owned, P["rowptr"], P["cols"])
    S, cap = L["n_slices"], int(L["adj_off"][-1])
    dm = np.ascontiguousarray(P["dofmap"], np.int32)
    rp = np.ascontiguousarray(P["rowptr"], np.int64)
    cols = L["cols"].copy()
    r = 20                                   # an interior-ish row: replace one real neighbour column
    mo, k = int(L["mat_off"][r >> 5]), int(rp[r + 1] - rp[r]) - 1
    assert cols[mo + k * 32 + (r & 31)] != r
    cols[mo + k * 32 + (r & 31)] = P.n_owned + P.n_ghost + 5
    adj_off, flags = np.zeros(S + 1, np.int64), np.zeros(2, np.int32)
    adjrot, walk = np.zeros(cap, np.uint32), np.zeros(cap, np.uint32)
    assert emusu.emu_setup_p1(C.c_int64(len(dm) // 4), _p(dm), P.n_owned, S, _p(rp), _p(L["mat_off"]), _p(cols),
                              0, C.c_int64(cap), _p(adj_off), _p(adjrot), _p(walk), _p(flags)) == 0
    assert flags[0] == 1


@pytest.mark.parametrize("ptype,order,dims,rank,nranks", [("poisson", 1, (5, 4, 6), 0, 1), ("poisson", 1, (1, 1, 1), 0, 1),
                                                          ("poisson", 1, (4, 3, 5), 1, 2), ("elasticity", 1, (3, 4, 3), 0, 1),
                                                          ("poisson", 2, (3, 2, 4), 0, 1), ("poisson", 2, (2, 2, 5), 1, 2),
                                                          ("poisson", 3, (2, 3, 2), 0, 1), ("poisson", 3, (2, 2, 4), 2, 3),
                                                          ("poisson", 1, (12, 11, 13), 0, 1), ("poisson", 3, (4, 4, 3), 0, 1)])
def test_device_pattern_source_equals_the_host_pattern(pt, emusu, ptype, order, dims, rank, nranks):
    """ptb_build_pattern's kernels (pairs -> per-row ascending union -> scan -> fill) against the
    pattern the host stand-in builds (common/intmaps.cpp build_pattern), P1-P3, with ghost columns."""
    P = pt.host.Problem(ptype, order, *dims, rank, nranks)
    dm = np.ascontiguousarray(P["dofmap"], np.int32)
    rowptr = np.full(P.n_owned + 1, -1, np.int64)
    cols = np.full(P.nnz, -1, np.int32)
    flags = np.full(3, -1, np.int32)
    rc = emusu.emu_build_pattern(C.c_int64(len(dm) // P.nd), P.nd, _p(dm), P.n_owned, C.c_int64(P.nnz), _p(rowptr),
                                 _p(cols), _p(flags))
    assert rc == 0 and flags.tolist() == [0, 0, 0]
    assert np.array_equal(rowptr, P["rowptr"])
    assert np.array_equal(cols, P["cols"])


# ---- device-side problem data: Dirichlet dofs and source terms (csrc/problem_data.cu) ---------------

PD_SRC = os.path.join(HERE, "emu", "emu_problem_data.cpp")


@pytest.fixture(scope="module")
def emupd():
    out = os.path.join(HERE, "emu", BUILD, "libemuproblemdata.so")
    deps = [PD_SRC] + [os.path.join(CSRC, f) for f in ("problem_data.cu", "kernels.h", "ctx.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        subprocess.run(["/usr/bin/g++", "-std=c++20", *CXXFLAGS, "-ffp-contract=off", "-fPIC", "-shared", "-w",
                        "-I", cuda_inc, "-I", CSRC, "-o", out, PD_SRC], check=True)
    return C.CDLL(out)


PROBLEM_DATA = [("poisson", 1, (5, 4, 6), 0, 1), ("poisson", 1, (1, 1, 1), 0, 1), ("poisson", 1, (4, 3, 5), 1, 2),
                ("elasticity", 1, (3, 4, 3), 0, 1), ("elasticity", 1, (2, 2, 5), 1, 2), ("poisson", 2, (3, 2, 4), 0, 1),
                ("poisson", 2, (2, 2, 5), 1, 2), ("poisson", 3, (2, 3, 2), 0, 1), ("poisson", 3, (2, 2, 4), 2, 3)]


@pytest.mark.parametrize("ptype,order,dims,rank,nranks", PROBLEM_DATA)
def test_bc_location_source_finds_the_reference_dofs(pt, emupd, ptype, order, dims, rank, nranks):
    """Facet marker + facet closure (locate_entities + locate_dofs_topological) on every local cell
    gives exactly the stand-in's Dirichlet dofs, owned and ghost, P1-P3, on partitions too."""
    P = pt.host.Problem(ptype, order, *dims, rank, nranks)
    x = np.asarray(P["x"]).reshape(-1, 3)
    xyz4 = np.zeros((len(x), 4))
    xyz4[:, :3] = x
    xd = np.ascontiguousarray(P["x_dofmap"], np.int32)
    dm = np.ascontiguousarray(P["dofmap"], np.int32)
    bc = np.zeros(P.n_owned + P.n_ghost, np.uint8)
    assert emupd.emu_locate_bc(C.c_int64(len(xd) // 4), pt.abi.PROBLEMS[ptype], order, P.nd, _p(xyz4), _p(xd), _p(dm),
                               _p(bc)) == 0
    assert np.array_equal(np.flatnonzero(bc), np.sort(P["bc_dofs"]))
    assert bc.max() <= 1


@pytest.mark.parametrize("ptype,order,dims,rank,nranks", PROBLEM_DATA)
def test_source_interpolation_source_equals_the_host_lambdas(pt, emupd, ptype, order, dims, rank, nranks):
    """f and g at the dof coordinates, bit for bit on the host (same libm, no contraction); for P1
    also through the padded by-dof vertex coordinates the context holds."""
    P = pt.host.Problem(ptype, order, *dims, rank, nranks)
    n = P.n_owned + P.n_ghost
    X = np.ascontiguousarray(P["dof_x"], np.float64)
    variants = [(X, 3)]
    if order == 1:
        X4 = np.zeros((n, 4))
        X4[:, :3] = X.reshape(-1, 3)
        variants.append((np.ascontiguousarray(X4.reshape(-1)), 4))
    for Xv, stride in variants:
        f = np.full(n * P.bs, np.nan)
        g = np.full(n, np.nan)
        assert emupd.emu_interpolate_source(C.c_int64(n), pt.abi.PROBLEMS[ptype], _p(Xv), stride, _p(f), _p(g)) == 0
        assert np.array_equal(f, P["f"])
        if ptype == "poisson":
            assert np.array_equal(g, P["g"])


@pytest.mark.parametrize("ptype,order,dims,rank,nranks", [("poisson", 1, (5, 4, 6), 0, 1), ("poisson", 1, (1, 1, 1), 0, 1),
                                                          ("poisson", 1, (33, 2, 1), 0, 1), ("poisson", 1, (4, 3, 5), 1, 2),
                                                          ("poisson", 1, (3, 3, 7), 2, 3), ("poisson", 1, (3, 3, 7), 0, 3),
                                                          ("poisson", 1, (12, 11, 13), 0, 1), ("poisson", 2, (3, 2, 4), 1, 2)])
def test_device_column_layout_source_equals_the_host_layout(pt, emusu, ptype, order, dims, rank, nranks):
    """SELL-32 offsets and padded columns, the compressed column indices (one delta per 32 rows /
    explicit lines) and the slice visiting order from the setup kernels equal layout.cpp's
    build_sell_layout / compress_columns / build_slice_order on single and partitioned boxes."""
    P = pt.host.Problem(ptype, order, *dims, rank, nranks)
    rp = np.ascontiguousarray(P["rowptr"], np.int64)
    cl = np.ascontiguousarray(P["cols"], np.int32)
    S = (P.n_owned + 31) // 32
    if order == 1:
        L = pt.abi.p1_layout(P["dofmap"], P.n_owned, rp, cl)
        mat_off_ref, cols_ref = L["mat_off"], L["cols"]
    else:
        L = pt.abi.pk_layout(P["dofmap"], P.nd, P.n_owned, rp, cl)
        mat_off_ref, cols_ref = L["mat_off"], L["cols"]
    cap = int(mat_off_ref[-1])
    cd_ref, xoff_ref, cx_ref = pt.abi.compressed_columns(P.n_owned, P.n_owned + P.n_ghost, rp, cl, cap)
    order_ref, ni_ref = pt.abi.slice_order(P.n_owned, rp, cl)
    mat_off, xoff = np.full(S + 1, -1, np.int64), np.full(S + 1, -1, np.int64)
    cols_sell, cdelta = np.full(cap, -7, np.int32), np.full(cap // 32, -7, np.int32)
    capx = int(xoff_ref[-1])
    colsx, so = np.full(max(capx, 1), -7, np.int32), np.full(S, -7, np.int32)
    ni = C.c_int32(-1)
    rc = emusu.emu_setup_columns(P.n_owned, C.c_int64(P.n_owned + P.n_ghost), S, _p(rp), _p(cl), C.c_int64(cap),
                                 C.c_int64(capx), _p(mat_off), _p(cols_sell), _p(cdelta), _p(xoff), _p(colsx), _p(so),
                                 C.byref(ni))
    assert rc == 0
    assert np.array_equal(mat_off, mat_off_ref) and np.array_equal(cols_sell, cols_ref)
    assert np.array_equal(cdelta, cd_ref) and np.array_equal(xoff, xoff_ref)
    assert np.array_equal(colsx[:capx], cx_ref[:capx])
    assert ni.value == ni_ref and np.array_equal(so, order_ref)


# ---- device-side mesh and P1 dofmap of the unit cube (csrc/box.cu) ----------------------------------

BX_SRC = os.path.join(HERE, "emu", "emu_box.cpp")


@pytest.fixture(scope="module")
def emubx():
    out = os.path.join(HERE, "emu", BUILD, "libemubox.so")
    deps = [BX_SRC] + [os.path.join(CSRC, f) for f in ("box.cu", "kernels.h", "ctx.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        subprocess.run(["/usr/bin/g++", "-std=c++20", *CXXFLAGS, "-ffp-contract=off", "-fPIC", "-shared", "-w",
                        "-I", cuda_inc, "-I", CSRC, "-o", out, BX_SRC], check=True)
    return C.CDLL(out)


def box_dims(nx, ny, nz, rank, nranks, order=1):
    """BoxDims as make_box_dims (box.cu) computes it: z-slabs, one ghost layer of cells below; the
    level stride of the numbering from the entity counts of one lattice level."""
    base, rem = divmod(nz, nranks)
    L0 = rank * base + min(rank, rem)
    L1 = L0 + base + (1 if rank < rem else 0)
    last = rank == nranks - 1
    ne, nf = order - 1, (order - 1) * (order - 2) // 2
    # plane block: vertices, x / y / xy edges, two faces per square; layer block: z / xz / yz / xyz
    # edges, six faces per cube interior + the rising faces over the x- and y-edges of the plane
    PS = ((nx + 1) * (ny + 1) + ne * (nx * (ny + 1) + (nx + 1) * ny + nx * ny) + nf * 2 * nx * ny)
    LS = (ne * ((nx + 1) * (ny + 1) + nx * (ny + 1) + (nx + 1) * ny + nx * ny)
          + nf * (6 * nx * ny + 2 * (nx + 1) * ny + 2 * nx * (ny + 1)))
    S = PS + LS
    l0 = L0 - 1 if rank > 0 else L0
    G0, G1 = L0 * S, (nz * S + PS if last else L1 * S)
    return np.array([nx, ny, nz, l0, L1, G0, G1, l0 * S, G1 if last else G1 + PS], np.int64)


@pytest.mark.parametrize("ptype,order,dims,rank,nranks", [("poisson", 1, (5, 4, 6), 0, 1), ("poisson", 1, (1, 1, 1), 0, 1),
                                                          ("poisson", 1, (4, 3, 5), 0, 2), ("poisson", 1, (4, 3, 5), 1, 2),
                                                          ("elasticity", 1, (3, 3, 7), 1, 3), ("poisson", 1, (3, 3, 7), 2, 3),
                                                          ("poisson", 1, (2, 7, 8), 5, 8), ("poisson", 2, (3, 2, 4), 0, 1),
                                                          ("poisson", 2, (2, 2, 5), 1, 2), ("poisson", 3, (2, 3, 2), 0, 1),
                                                          ("poisson", 3, (2, 2, 4), 2, 3), ("poisson", 3, (1, 1, 3), 1, 3)])
def test_box_generator_source_equals_the_host_mesh_and_dofmap(pt, emubx, ptype, order, dims, rank, nranks):
    """Vertices, cell -> vertex map, Lagrange dofmap (P1-P3), dof coordinates and the dof -> vertex
    inverse from the generator kernels equal the host stand-in's arrays bit for bit on every rank of
    a z-slab partition."""
    P = pt.host.Problem(ptype, order, *dims, rank, nranks)
    B = box_dims(*dims, rank, nranks, order)
    x_ref = np.asarray(P["x"])
    nv, nc = len(x_ref) // 3, len(P["x_dofmap"]) // 4
    n = P.n_owned + P.n_ghost
    assert int(B[6] - B[5]) == P.n_owned and int((B[5] - B[7]) + (B[8] - B[6])) == P.n_ghost
    xyz3, xyz4 = np.full(nv * 3, np.nan), np.full(nv * 4, np.nan)
    dv = np.full(n, -1, np.int32)
    xd, dm = np.full(nc * 4, -5, np.int32), np.full(nc * P.nd, -5, np.int32)
    dof_x = np.full(n * 3, np.nan)
    flags = np.full(1, -1, np.int32)
    dims_out = np.full(9, -1, np.int64)
    assert emubx.emu_create_box(C.c_int64(dims[0]), C.c_int64(dims[1]), C.c_int64(dims[2]), rank, nranks, order,
                                _p(dims_out), _p(xyz3), _p(xyz4), _p(dv), _p(xd), _p(dm), _p(dof_x), _p(flags)) == 0
    assert flags[0] == 0
    assert np.array_equal(dims_out, B)          # the product's slab / range computation (make_box_dims)
    assert np.array_equal(xyz3, x_ref)
    assert np.array_equal(xyz4.reshape(-1, 4)[:, :3].reshape(-1), x_ref) and np.all(xyz4.reshape(-1, 4)[:, 3] == 0)
    assert np.array_equal(xd, P["x_dofmap"]) and np.array_equal(dm, P["dofmap"])
    assert np.array_equal(dof_x, P["dof_x"])
    # dof -> vertex is the inverse of the vertex dofs of the cells, -1 for edge and face dofs
    inv = np.full(n, -1, np.int32)
    inv[np.asarray(P["dofmap"]).reshape(-1, P.nd)[:, :4].reshape(-1)] = np.asarray(P["x_dofmap"])
    assert np.array_equal(dv, inv)


@pytest.mark.parametrize("ptype,order,dims,rank,nranks", [("poisson", 1, (5, 4, 6), 0, 1), ("poisson", 1, (4, 3, 5), 1, 2),
                                                          ("poisson", 2, (3, 2, 4), 0, 1), ("poisson", 3, (2, 3, 2), 1, 2)])
def test_facet_rows_from_gathered_dofmap_rows(pt, emubx, ptype, order, dims, rank, nranks):
    """A context whose dofmap lives on the device downloads the rows of the exterior facets' cells
    only (gather kernel) and builds the same boundary-facet row lists from them."""
    P = pt.host.Problem(ptype, order, *dims, rank, nranks)
    fc = np.ascontiguousarray(P["facet_cells"], np.int32)
    dm = np.ascontiguousarray(P["dofmap"], np.int32)
    rows = np.full(len(fc) * P.nd, -3, np.int32)
    assert emubx.emu_gather_dofmap_rows(C.c_int64(len(fc)), P.nd, _p(fc), _p(dm), _p(rows)) == 0
    assert np.array_equal(rows, dm.reshape(-1, P.nd)[fc].reshape(-1))
    ref = pt.abi.facet_rows(fc, P["facet_local"], dm, P.nd, order, P.n_owned)
    got = pt.abi.facet_rows(fc, P["facet_local"], rows, P.nd, order, P.n_owned, gathered=True)
    assert all(np.array_equal(a, b) for a, b in zip(ref, got)) and len(ref[0]) > 0


@pytest.mark.parametrize("order,dims,rank,nranks", [(2, (3, 2, 4), 0, 1), (2, (2, 2, 5), 1, 2), (3, (2, 3, 2), 0, 1),
                                                    (3, (2, 2, 4), 2, 3)])
def test_device_setup_source_builds_the_p2_p3_slot_words(pt, emusu, order, dims, rank, nranks):
    """adj_off, the pair words and the packed 8-bit slot offsets of the P2/P3 assembly kernels from
    setup_adj_pk equal layout.cpp's build_sell_layout word for word (pairs reversed before the sort)."""
    P = pt.host.Problem("poisson", order, *dims, rank, nranks)
    L = pt.abi.pk_layout(P["dofmap"], P.nd, P.n_owned, P["rowptr"], P["cols"])
    assert L["so_bits"] == 8 and L["so_words"] == (P.nd + 3) // 4
    S, cap = L["n_slices"], int(L["adj_off"][-1])
    dm = np.ascontiguousarray(P["dofmap"], np.int32)
    rp = np.ascontiguousarray(P["rowptr"], np.int64)
    adj_off = np.full(S + 1, -1, np.int64)
    adj = np.full(cap, 0xDEADBEEF, np.uint32)
    adjso = np.full(cap * L["so_words"], 0xDEADBEEF, np.uint32)
    flags = np.full(2, -1, np.int32)
    rc = emusu.emu_setup_pk(C.c_int64(len(dm) // P.nd), P.nd, _p(dm), P.n_owned, S, _p(rp), _p(L["mat_off"]),
                            _p(L["cols"]), C.c_int64(cap), _p(adj_off), _p(adj), _p(adjso), _p(flags))
    assert rc == 0 and flags.tolist() == [0, 0]
    assert np.array_equal(adj_off, L["adj_off"])
    assert np.array_equal(adj, L["adj"])
    assert np.array_equal(adjso, L["adjso"])


@pytest.mark.parametrize("order", [1, 2])
def test_device_pattern_column_and_map_builders_on_an_unstructured_mesh(pt, emusu, order):
    """ptb_build_pattern, gpu_setup_columns and gpu_setup_pk (setup.cu, emulated) on the Delaunay mesh:
    the pattern equals the sorted union of the cells' dofs, the column side and the P2 pair / slot
    words equal the host build word for word. (Nothing here is translation invariant: every column
    index of the scalar operator goes explicit.)"""
    P = _Unstructured("poisson", 90, 7, order=order)
    dm = np.ascontiguousarray(P["dofmap"], np.int32)
    rp, cl = np.ascontiguousarray(P["rowptr"], np.int64), np.ascontiguousarray(P["cols"], np.int32)
    nnz = int(rp[-1])
    rowptr, cols, flags3 = np.full(P.n_owned + 1, -1, np.int64), np.full(nnz, -1, np.int32), np.full(3, -1, np.int32)
    rc = emusu.emu_build_pattern(C.c_int64(P.n_cells), P.nd, _p(dm), P.n_owned, C.c_int64(nnz), _p(rowptr), _p(cols),
                                 _p(flags3))
    if flags3[2] != 0:
        pytest.skip("a row beyond the device pattern kernel's capacity: the product falls back to the host build")
    assert rc == 0 and np.array_equal(rowptr, rp) and np.array_equal(cols, cl)
    L = (pt.abi.p1_layout if order == 1 else lambda d, n, r, c: pt.abi.pk_layout(d, P.nd, n, r, c))(dm, P.n_owned, rp, cl)
    S, cap = L["n_slices"], int(L["mat_off"][-1])
    cd_ref, xoff_ref, cx_ref = pt.abi.compressed_columns(P.n_owned, P.n_owned, rp, cl, cap)
    order_ref, ni_ref = pt.abi.slice_order(P.n_owned, rp, cl)
    mat_off, xoff = np.full(S + 1, -1, np.int64), np.full(S + 1, -1, np.int64)
    cols_sell, cdelta = np.full(cap, -7, np.int32), np.full(cap // 32, -7, np.int32)
    capx = int(xoff_ref[-1])
    colsx, so = np.full(max(capx, 1), -7, np.int32), np.full(S, -7, np.int32)
    ni = C.c_int32(-1)
    assert emusu.emu_setup_columns(P.n_owned, C.c_int64(P.n_owned), S, _p(rp), _p(cl), C.c_int64(cap), C.c_int64(capx),
                                   _p(mat_off), _p(cols_sell), _p(cdelta), _p(xoff), _p(colsx), _p(so), C.byref(ni)) == 0
    assert np.array_equal(mat_off, L["mat_off"]) and np.array_equal(cols_sell, L["cols"])
    assert np.array_equal(cdelta, cd_ref) and np.array_equal(xoff, xoff_ref) and np.array_equal(colsx[:capx], cx_ref[:capx])
    assert ni.value == ni_ref and np.array_equal(so, order_ref)
    assert capx > 0.9 * cap                       # (almost) no translation-invariant column on this mesh
    if order == 2:
        capa = int(L["adj_off"][-1])
        adj_off, adj = np.full(S + 1, -1, np.int64), np.full(capa, 0xDEADBEEF, np.uint32)
        adjso, flags = np.full(capa * L["so_words"], 0xDEADBEEF, np.uint32), np.full(2, -1, np.int32)
        assert L["so_bits"] == 8
        assert emusu.emu_setup_pk(C.c_int64(P.n_cells), P.nd, _p(dm), P.n_owned, S, _p(rp), _p(L["mat_off"]), _p(L["cols"]),
                                  C.c_int64(capa), _p(adj_off), _p(adj), _p(adjso), _p(flags)) == 0
        assert flags.tolist() == [0, 0] and np.array_equal(adj_off, L["adj_off"])
        assert np.array_equal(adj, L["adj"]) and np.array_equal(adjso, L["adjso"])


@pytest.mark.parametrize("n,scale", [(1, 1), (1000, 32), (8192, 1), (8193, 1), (16384, 32), (20001, 1)])
def test_device_scan_source_both_routes(emusu, n, scale):
    """The prefix sum of the setup kernels: one CTA up to a tile (8192 elements), three tile passes
    beyond; exclusive, scaled, total at out[n]."""
    rng = np.random.default_rng(n)
    a = rng.integers(0, 50, size=n).astype(np.uint64)
    out = np.full(n + 1, -1, np.int64)
    assert emusu.emu_scan_only(C.c_int64(n), _p(a), C.c_int64(scale), _p(out)) == 0
    ref = np.concatenate([[0], np.cumsum(a.astype(np.int64) * scale)])
    assert np.array_equal(out, ref)
