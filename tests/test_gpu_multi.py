"""Partitioned (multi-rank) CUDA path.

* test_partitioned_assembly_...: every rank's slab is assembled on ONE GPU (no communication is
  needed for assembly: ghost-cell layer) and must reproduce the serial rows bit for bit.
* test_two_rank_solve_...: needs >= 2 GPUs (skipped otherwise): one process per GPU, NCCL halo +
  all-reduce inside ptb_cg_solve, compared with the single-partition oracle.
"""
import importlib
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("ptype,order,dims,world",
                         [("poisson", 1, (9, 8, 10), 3), ("elasticity", 1, (6, 5, 8), 2),
                          ("poisson", 2, (4, 5, 6), 2), ("poisson", 3, (3, 3, 5), 2)])
def test_partitioned_assembly_is_partition_independent(pt, ptype, order, dims, world):
    S = pt.host.Problem(ptype, order, *dims)
    ctx = pt.abi.Context(0)
    ctx.set_problem(S)
    ctx.assemble_matrix()
    ctx.assemble_vector()
    A_s, b_s = ctx.matrix_values().copy(), ctx.rhs().copy()
    bs = S.bs
    for rank in range(world):
        P = pt.host.Problem(ptype, order, *dims, rank, world)
        ctx.set_problem(P)
        ctx.assemble_matrix()
        ctx.assemble_vector()
        A, b = ctx.matrix_values(), ctx.rhs()
        off, n = P.global_offset, P.n_owned
        assert np.array_equal(b, b_s[off * bs:(off + n) * bs])
        l2g = np.concatenate([np.arange(off, off + n), P["ghost_global"]])
        rp, cl = P["rowptr"], P["cols"]
        for r in range(n):
            g = l2g[cl[rp[r]:rp[r + 1]]]
            o = np.argsort(g)
            srow = slice(S["rowptr"][off + r], S["rowptr"][off + r + 1])
            assert np.array_equal(g[o], S["cols"][srow])
            blk, blk_s = A.reshape(-1, bs * bs)[rp[r]:rp[r + 1]][o], A_s.reshape(-1, bs * bs)[srow]
            if ptype == "elasticity" and order == 1:
                # edge-ring kernel (assemble_ring.cu): off-diagonal blocks bit for bit; the diagonal block
                # is -sum of the row's blocks in local column order (ghost columns last): a few ulp
                dg = S["cols"][srow] == off + r
                assert np.array_equal(blk[~dg], blk_s[~dg])
                assert np.abs(blk[dg] - blk_s[dg]).max() <= 1e-14 * np.abs(blk_s[dg]).max()
            else:
                assert np.array_equal(blk, blk_s)
    ctx.close()


@pytest.mark.parametrize("ptype,order,dims,world",
                         [("poisson", 1, (9, 8, 10), 3), ("elasticity", 1, (6, 5, 8), 2),
                          ("poisson", 2, (4, 5, 6), 2), ("poisson", 3, (3, 3, 5), 2)])
def test_device_generated_slabs_are_partition_independent(pt, monkeypatch, ptype, order, dims, world):
    """The same check with every rank's slab generated on the device (ptb_create_box with rank > 0:
    ghost layer below, ghost plane above; pattern, layouts and maps device-built with
    PTB_GPU_SETUP=1): rows and right-hand side equal the serial host-built ones bit for bit, except
    that f and g come from the device's exp / sin (<= 4 ulp), so b is compared to 1e-14."""
    S = pt.host.Problem(ptype, order, *dims)
    monkeypatch.setenv("PTB_GPU_SETUP", "1")
    ctx = pt.abi.Context(0)
    ctx.set_problem(S)
    ctx.assemble_matrix()
    ctx.assemble_vector()
    A_s, b_s = ctx.matrix_values().copy(), ctx.rhs().copy()
    bs = S.bs
    for rank in range(world):
        P = pt.host.Problem(ptype, order, *dims, rank, world)
        ctx.set_problem_on_device(P)
        assert np.array_equal(ctx.dofmap(), P["dofmap"]) and np.array_equal(ctx.mesh()[0], P["x"])
        rp, cl = ctx.pattern()
        assert np.array_equal(rp, P["rowptr"]) and np.array_equal(cl, P["cols"])
        ctx.assemble_matrix()
        ctx.assemble_vector()
        A, b = ctx.matrix_values(), ctx.rhs()
        off, n = P.global_offset, P.n_owned
        assert np.abs(b - b_s[off * bs:(off + n) * bs]).max() <= 1e-14 * np.abs(b_s).max()
        l2g = np.concatenate([np.arange(off, off + n), P["ghost_global"]])
        for r in range(n):
            g = l2g[cl[rp[r]:rp[r + 1]]]
            o = np.argsort(g)
            srow = slice(S["rowptr"][off + r], S["rowptr"][off + r + 1])
            assert np.array_equal(g[o], S["cols"][srow])
            blk, blk_s = A.reshape(-1, bs * bs)[rp[r]:rp[r + 1]][o], A_s.reshape(-1, bs * bs)[srow]
            if ptype == "elasticity" and order == 1:   # edge-ring kernel: see the test above
                dg = S["cols"][srow] == off + r
                assert np.array_equal(blk[~dg], blk_s[~dg])
                assert np.abs(blk[dg] - blk_s[dg]).max() <= 1e-14 * np.abs(blk_s[dg]).max()
            else:
                assert np.array_equal(blk, blk_s)
    ctx.close()


def _worker(rank, world, port, ptype, dims, comm, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pt = importlib.import_module("performance-test_b200")
        ctx = pt.abi.Context(rank)
        P = pt.host.Problem(ptype, 1, *dims, rank, world)
        if os.environ.get("PTB_TEST_DEVICE_SETUP") == "1":  # opt-in: the slab generated on the device
            ctx.set_problem_on_device(P)
        else:
            ctx.set_problem(P)
        if comm == "nccl":
            pt.dist.init_nccl(ctx, pt.abi, dist, rank, world)
        else:
            pt.dist.connect_peers(ctx, P, dist, rank, world)
        ctx.assemble_matrix()
        ctx.assemble_vector()
        k, rel = ctx.cg_solve(kmax=5000, rtol=1e-8, precond="jacobi")
        if ptype == "poisson":  # the matrix-free action must take the same iterates
            ctx.set_operator_mode("matrix_free")
            k_mf, rel_mf = ctx.cg_solve(kmax=5000, rtol=1e-8, precond="jacobi")
            assert abs(k_mf - k) <= 1 and rel_mf < 1e-8
            ctx.set_operator_mode("assembled")
            k, rel = ctx.cg_solve(kmax=5000, rtol=1e-8, precond="jacobi")
        x = ctx.solution()
        nrm = ctx.solution_norm()
        rng = np.random.default_rng(7)
        pg = rng.standard_normal(P.n_global * P.bs)  # same on every rank
        pl = np.zeros((P.n_owned + P.n_ghost) * P.bs)
        pl[: P.n_owned * P.bs] = pg[P.global_offset * P.bs:(P.global_offset + P.n_owned) * P.bs]
        y = ctx.apply_operator(pl)  # ghosts of p are filled by the halo exchange
        out.put((rank, k, rel, P.global_offset, P.n_owned, x, nrm, y, np.array(P["ghost_global"])))
        dist.barrier()  # nobody frees device memory a peer may still map
        ctx.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("comm", ["nccl", "peer"])
@pytest.mark.parametrize("ptype,dims", [("poisson", (12, 11, 14)), ("elasticity", (7, 8, 9)),
                                        ("poisson", (24, 23, 30))])
def test_two_rank_solve_matches_oracle(pt, oracle, ptype, dims, comm):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    ctxm = mp.get_context("spawn")
    out = ctxm.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctxm.Process(target=_worker, args=(r, world, port, ptype, dims, comm, out))
             for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = sorted([out.get(timeout=150) for _ in range(world)], key=lambda t: t[0])
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:
        for p in procs:  # never leave a spinning rank behind (a hung peer kernel would block the box)
            if p.is_alive():
                p.kill()
    S = pt.host.Problem(ptype, 1, *dims)
    bs = S.bs
    A_s, b_s = oracle.assemble_matrix(S), oracle.assemble_vector(S)
    x_s, k_s, rel_s = oracle.cg(bs, S.n_owned, S["rowptr"], S["cols"], A_s, b_s, kmax=5000,
                                rtol=1e-8, precond="jacobi")
    rng = np.random.default_rng(7)
    pg = rng.standard_normal(S.n_global * bs)
    y_s = oracle.spmv(bs, S.n_owned, S["rowptr"], S["cols"], A_s, pg)
    xg = np.zeros(S.n_owned * bs)
    for rank, k, rel, off, n, x, nrm, y, gg in res:
        assert abs(k - k_s) <= 1 and rel < 1e-8
        xg[off * bs:(off + n) * bs] = x[: n * bs]
        assert np.abs(y - y_s[off * bs:(off + n) * bs]).max() <= 1e-12 * np.abs(y_s).max()
    for rank, k, rel, off, n, x, nrm, y, gg in res:
        assert np.array_equal(x.reshape(-1, bs)[n:], xg.reshape(-1, bs)[gg])  # ghosts current
        assert nrm == pytest.approx(np.linalg.norm(xg), rel=1e-12)
    assert np.linalg.norm(xg - x_s) <= 1e-6 * np.linalg.norm(x_s)
