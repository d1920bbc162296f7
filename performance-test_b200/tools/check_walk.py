"""GPU smoke + A/B of the opt-in star-walk assembly kernel (PTB_ASM_WALK=1) against the default
kernel. Small parity case first, then one timing case; results are appended to
gpurun_out/walk_check.json after every stage so that a cut-off run still leaves evidence.

    python performance-test_b200/tools/check_walk.py [ndofs_for_timing]
"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
pt = importlib.import_module("performance-test_b200")
OUT = os.path.join(ROOT, "gpurun_out", os.environ.get("WALK_CHECK_OUT", "walk_check.json"))
res = {}


def dump():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res), flush=True)


def assemble(P, walk):
    if walk:
        os.environ["PTB_ASM_WALK"] = "1"
    else:
        os.environ.pop("PTB_ASM_WALK", None)
    c = pt.abi.Context(0)
    c.set_problem(P)
    c.assemble_matrix()
    return c


def ab(ndofs):
    """One walk context; launch-time knobs (PTB_WALK_WARPS, PTB_WALK_PREFETCH) varied in place."""
    nx, ny, nz, r = pt.host.cube_sizing(ndofs, True, 1, 1, 1)
    f = 2 ** r
    P = pt.host.Problem("poisson", 1, nx * f, ny * f, nz * f)
    c = assemble(P, True)
    res["ab_case"] = {"n_owned": P.n_owned, "nnz": P.nnz}
    for warps in (2, 1, 4, 7):
        for pf in (0, 1):
            os.environ["PTB_WALK_WARPS"], os.environ["PTB_WALK_PREFETCH"] = str(warps), str(pf)
            res[f"ms_warps{warps}_pf{pf}"] = c.time_kernel(pt.abi.KERNEL_ASSEMBLE_MATRIX, 5)
            dump()
    c.close()


def ab2(ndofs):
    """Matrix and vector assembly: the default kernels against the generation they replaced
    (PTB_VEC_GWALK=0 staged-star vector kernel, PTB_ASM_WALK3=0 / PTB_ASM_WALK=0 cell-order matrix
    kernels), one context per variant."""
    for ptype, dpn in (("poisson", 1), ("elasticity", 3)):
        nx, ny, nz, r = pt.host.cube_sizing(ndofs, True, dpn, 1, 1)
        f = 2 ** r
        P = pt.host.Problem(ptype, 1, nx * f, ny * f, nz * f)
        ref = None
        variants = [("default", {}), ("staged_vector", {"PTB_VEC_GWALK": "0"})]
        if ptype == "elasticity":
            variants.append(("cellorder", {"PTB_ASM_WALK3": "0"}))
        else:
            variants.append(("cellorder", {"PTB_ASM_WALK": "0"}))
        for name, env in variants:
            for k in ("PTB_VEC_GWALK", "PTB_ASM_WALK3", "PTB_ASM_WALK"):
                os.environ.pop(k, None)
            os.environ.update(env)
            c = pt.abi.Context(0)
            c.set_problem(P)
            c.assemble_matrix()
            c.assemble_vector()
            a, b = c.matrix_values(), c.rhs()
            if ref is None:
                ref = (a, b)
            key = f"{ptype}_{name}"
            res[key] = {"matrix_ms": c.time_kernel(pt.abi.KERNEL_ASSEMBLE_MATRIX, 5),
                        "vector_ms": c.time_kernel(pt.abi.KERNEL_ASSEMBLE_VECTOR, 5),
                        "matrix_maxdiff": float(np.abs(a - ref[0]).max() / np.abs(ref[0]).max()),
                        "vector_maxdiff": float(np.abs(b - ref[1]).max() / np.abs(ref[1]).max()),
                        "n_owned": P.n_owned, "nnz": P.nnz}
            c.close()
            dump()


def abpk(ndofs):
    """P2 / P3 matrix assembly: one launch over all slices against the row-length bins."""
    for order in (2, 3):
        nx, ny, nz, r = pt.host.cube_sizing(ndofs, True, 1, order, 1)
        f = 2 ** r
        P = pt.host.Problem("poisson", order, nx * f, ny * f, nz * f)
        c = pt.abi.Context(0)
        c.set_problem(P)
        ref = None
        for bins in ("0", "1"):
            os.environ["PTB_PK_BINS"] = bins
            c.assemble_matrix()
            a = c.matrix_values()
            ref = a if ref is None else ref
            res[f"p{order}_bins{bins}"] = {"matrix_ms": c.time_kernel(pt.abi.KERNEL_ASSEMBLE_MATRIX, 3),
                                           "maxdiff": float(np.abs(a - ref).max() / np.abs(ref).max()),
                                           "n_owned": P.n_owned, "nnz": P.nnz}
            dump()
        c.close()


def abmf(ndofs):
    """Matrix-free CG (cgpoisson's action) on Poisson P1 next to the assembled operator;
    DOF-iterations/s from the device time of the solve stage."""
    nx, ny, nz, r = pt.host.cube_sizing(ndofs, True, 1, 1, 1)
    f = 2 ** r
    P = pt.host.Problem("poisson", 1, nx * f, ny * f, nz * f)
    c = pt.abi.Context(0)
    c.set_problem(P)
    c.assemble_matrix()
    c.assemble_vector()
    for mode in ("assembled", "matrix_free"):
        c.set_operator_mode(mode)
        k, rel = c.cg_solve(kmax=200, rtol=1e-30)
        ms = c.stage_ms(pt.abi.STAGE_SOLVE)
        res[f"cg_{mode}"] = {"iterations": k, "solve_ms": ms, "gdof_it_per_s": k * P.n_owned / ms / 1e6}
        dump()
    c.close()


def absetup(ndofs):
    """ptb_set_* wall time with the P1 assembly maps built on the host and on the device
    (PTB_GPU_SETUP=1, csrc/setup.cu); the assembled matrices must be identical."""
    nx, ny, nz, r = pt.host.cube_sizing(ndofs, True, 1, 1, 1)
    f = 2 ** r
    P = pt.host.Problem("poisson", 1, nx * f, ny * f, nz * f)
    res["setup_case"] = {"n_owned": P.n_owned, "nnz": P.nnz}
    sums = {}
    for dev in ("0", "1", "0", "1"):
        os.environ["PTB_GPU_SETUP"] = dev
        c = pt.abi.Context(0)
        t0 = time.perf_counter()
        c.set_problem(P)
        res[f"set_problem_s_gpu_setup{dev}"] = min(res.get(f"set_problem_s_gpu_setup{dev}", 1e9),
                                                  time.perf_counter() - t0)
        res[f"built_on_device_gpu_setup{dev}"] = c.p1_maps()["built_on_device"]
        c.assemble_matrix()
        sums[dev] = float(np.abs(c.matrix_values()).sum())
        c.close()
        dump()
    res["setup_checksums_equal"] = sums["0"] == sums["1"]
    # the pattern itself (the reference's create_matrix step) on the device, then the maps
    os.environ["PTB_GPU_SETUP"] = "1"
    c = pt.abi.Context(0)
    t0 = time.perf_counter()
    c.set_problem(P, build_pattern=True)
    res["set_problem_s_device_pattern_and_maps"] = time.perf_counter() - t0
    rp, cl = c.pattern()
    res["device_pattern_equal"] = bool(np.array_equal(rp, P["rowptr"]) and np.array_equal(cl, P["cols"]))
    c.close()
    dump()
    # everything generated on the device: mesh, dofmap, pattern, layouts, maps, Dirichlet dofs, sources
    c = pt.abi.Context(0)
    t0 = time.perf_counter()
    c.set_problem_on_device(P)
    res["set_problem_on_device_s"] = time.perf_counter() - t0
    c.assemble_matrix()
    res["on_device_checksum_equal"] = float(np.abs(c.matrix_values()).sum()) == sums["1"]
    c.close()
    dump()


def ncu_target(ndofs):
    nx, ny, nz, r = pt.host.cube_sizing(ndofs, True, 1, 1, 1)
    f = 2 ** r
    P = pt.host.Problem("poisson", 1, nx * f, ny * f, nz * f)
    assemble(P, True).close()


def ncu_target2(ndofs):
    """Matrix + vector assembly once with whatever switches the environment carries."""
    nx, ny, nz, r = pt.host.cube_sizing(ndofs, True, 1, 1, 1)
    f = 2 ** r
    P = pt.host.Problem("poisson", 1, nx * f, ny * f, nz * f)
    c = pt.abi.Context(0)
    c.set_problem(P)
    c.assemble_matrix()
    c.assemble_vector()
    c.close()


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "ncu2":
        return ncu_target2(int(sys.argv[2]))
    if len(sys.argv) > 2 and sys.argv[1] == "ab":
        return ab(int(sys.argv[2]))
    if len(sys.argv) > 2 and sys.argv[1] == "ab2":
        return ab2(int(sys.argv[2]))
    if len(sys.argv) > 2 and sys.argv[1] == "abmf":
        return abmf(int(sys.argv[2]))
    if len(sys.argv) > 2 and sys.argv[1] == "absetup":
        return absetup(int(sys.argv[2]))
    if len(sys.argv) > 2 and sys.argv[1] == "abpk":
        return abpk(int(sys.argv[2]))
    if len(sys.argv) > 2 and sys.argv[1] == "ncu":
        return ncu_target(int(sys.argv[2]))
    t0 = time.time()
    for dims in [(5, 4, 6), (16, 15, 17)]:
        P = pt.host.Problem("poisson", 1, *dims)
        c0, c1 = assemble(P, False), assemble(P, True)
        a0, a1 = c0.matrix_values(), c1.matrix_values()
        d0, d1 = c0.diagonal_inverse(), c1.diagonal_inverse()
        scale = np.abs(a0).max()
        res[f"parity_{dims}"] = {"max_abs_diff_over_max": float(np.abs(a0 - a1).max() / scale),
                                 "dinv_rel": float(np.abs(d0 / d1 - 1).max()),
                                 "nan": bool(np.isnan(a1).any())}
        c0.close(), c1.close()
        dump()
    ndofs = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
    nx, ny, nz, r = pt.host.cube_sizing(ndofs, True, 1, 1, 1)
    f = 2 ** r
    P = pt.host.Problem("poisson", 1, nx * f, ny * f, nz * f)
    res["timing_case"] = {"n_owned": P.n_owned, "nnz": P.nnz, "setup_s": time.time() - t0}
    dump()
    for walk in (False, True):
        c = assemble(P, walk)
        ms = c.time_kernel(pt.abi.KERNEL_ASSEMBLE_MATRIX, 5)
        res[f"assemble_matrix_ms_walk{int(walk)}"] = ms
        res[f"gnnz_per_s_walk{int(walk)}"] = P.nnz / ms / 1e6
        if walk:
            res["checksum_walk"] = float(np.abs(c.matrix_values()).sum())
        else:
            res["checksum_default"] = float(np.abs(c.matrix_values()).sum())
        c.close()
        dump()


if __name__ == "__main__":
    main()
