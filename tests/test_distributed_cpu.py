"""N > 1 host logic on CPU: world_size-2/3 gloo runs of the partitioned problem.

Each rank builds its z-slab (host stand-in), assembles its owned rows with the oracle, and runs the
cg.h loop with the halo exchange (Scatterer-style index lists, cgpoisson_problem.cpp:212-229) and
the global dot products (MPI_Allreduce in la::inner_product) done over torch.distributed/gloo.
The result must match the single-partition oracle: same matrix rows, iteration count +-1, same
solution. This pins ownership, ghost numbering and the halo lists that the CUDA path consumes.
"""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _halo_forward(dist, P, v, bs):
    """owner -> ghost update of v ([owned*bs | ghost*bs]) with the problem's halo lists."""
    reqs, recvs = [], []
    import torch
    sd, rd = P["send_displ"], P["recv_displ"]
    li, ri = P["local_indices"], P["remote_indices"]
    vb = v.reshape(-1, bs)
    for i, nbr in enumerate(P["nbr_ranks"]):
        out = torch.from_numpy(np.ascontiguousarray(vb[li[sd[i]:sd[i + 1]]]))
        buf = torch.empty((rd[i + 1] - rd[i], bs), dtype=torch.float64)
        reqs.append(dist.isend(out, int(nbr)))
        reqs.append(dist.irecv(buf, int(nbr)))
        recvs.append((i, buf))
    for r in reqs:
        r.wait()
    for i, buf in recvs:
        vb[ri[rd[i]:rd[i + 1]]] = buf.numpy()


def _worker(rank, world, port, ptype, order, dims, precond, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pt = importlib.import_module("performance-test_b200")
        import oracle
        P = pt.host.Problem(ptype, order, *dims, rank, world)
        bs, n = P.bs, P.n_owned * P.bs
        A = oracle.assemble_matrix(P)
        b = oracle.assemble_vector(P)
        dinv = np.ones(n)
        if precond == "jacobi":
            rows = np.repeat(np.arange(P.n_owned), np.diff(P["rowptr"]))
            dinv = 1.0 / np.einsum("kii->ki", A.reshape(-1, bs, bs)[P["cols"] == rows]).reshape(-1)

        def gsum(v):
            t = torch.tensor([v], dtype=torch.float64)
            dist.all_reduce(t)
            return float(t[0])

        def action(p_local):
            _halo_forward(dist, P, p_local, bs)
            return oracle.spmv(bs, P.n_owned, P["rowptr"], P["cols"], A, p_local)

        nl = (P.n_owned + P.n_ghost) * bs
        x, p = np.zeros(nl), np.zeros(nl)
        r = b - action(x)
        p[:n] = dinv * r
        rnorm0 = rnorm = gsum(r @ r)
        rz = gsum(r @ (dinv * r))
        k = 0
        while k < 5000:
            k += 1
            y = action(p)
            alpha = rz / gsum(p[:n] @ y)
            x[:n] += alpha * p[:n]
            r -= alpha * y
            rnorm = gsum(r @ r)
            rz_new = gsum(r @ (dinv * r))
            beta, rz = rz_new / rz, rz_new
            if rnorm / rnorm0 < 1e-16:
                break
            p[:n] = beta * p[:n] + dinv * r
        _halo_forward(dist, P, x, bs)
        out.put((rank, k, P.global_offset, P.n_owned, x.copy(), A.copy(), b.copy(),
                 np.array(P["ghost_global"]), np.array(P["rowptr"]), np.array(P["cols"])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ptype,order,dims,world,precond",
                         [("poisson", 1, (6, 5, 8), 2, "jacobi"),
                          ("elasticity", 1, (4, 4, 7), 2, "jacobi"),
                          ("poisson", 1, (5, 4, 9), 3, "none")])
def test_partitioned_oracle_cg_matches_serial(ptype, order, dims, world, precond):
    import torch.multiprocessing as mp
    pt = importlib.import_module("performance-test_b200")
    import oracle
    ctxm = mp.get_context("spawn")
    out = ctxm.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctxm.Process(target=_worker, args=(r, world, port, ptype, order, dims, precond, out))
             for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([out.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    S = pt.host.Problem(ptype, order, *dims)
    bs = S.bs
    A_s, b_s = oracle.assemble_matrix(S), oracle.assemble_vector(S)
    x_s, k_s, _ = oracle.cg(bs, S.n_owned, S["rowptr"], S["cols"], A_s, b_s, kmax=5000, rtol=1e-8,
                            precond=precond)
    x_glob = np.zeros(S.n_owned * bs)
    for rank, k, off, n_owned, x, A, b, gg, rp, cl in res:
        assert abs(k - k_s) <= 1
        x_glob[off * bs:(off + n_owned) * bs] = x[: n_owned * bs]
        # ghost values equal the owners' values
        l2g = np.concatenate([np.arange(off, off + n_owned), gg])
        # owned rows and RHS are bit-identical to the serial ones (same cells, same order)
        np.testing.assert_array_equal(b, b_s[off * bs:(off + n_owned) * bs])
        for r in range(n_owned):
            gcols = l2g[cl[rp[r]:rp[r + 1]]]
            order_ = np.argsort(gcols)
            srow = slice(S["rowptr"][off + r], S["rowptr"][off + r + 1])
            assert np.array_equal(gcols[order_], S["cols"][srow])
            assert np.array_equal(A.reshape(-1, bs * bs)[rp[r]:rp[r + 1]][order_],
                                  A_s.reshape(-1, bs * bs)[srow])
    for rank, k, off, n_owned, x, A, b, gg, rp, cl in res:
        np.testing.assert_array_equal(x.reshape(-1, bs)[n_owned:], x_glob.reshape(-1, bs)[gg])
    assert np.linalg.norm(x_glob - x_s) <= 1e-7 * np.linalg.norm(x_s)
