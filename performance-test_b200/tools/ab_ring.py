"""GPU A/B of the elasticity P1 matrix kernels: column-major along the edge rings
(assemble_matrix_p1_ring3, PTB_RING_WARPS = 4 / 2 / 1) against the star walk (PTB_ASM_RING=0) and the
first-generation kernel (PTB_ASM_RING=0 PTB_ASM_WALK3=0); the matrices are compared entry by entry
(bound: 1e-12 of the row's diagonal). Results go to gpurun_out/ring_ab.json after every stage.

    python performance-test_b200/tools/ab_ring.py [ndofs ...]
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
OUT = os.path.join(ROOT, "gpurun_out", os.environ.get("RING_AB_OUT", "ring_ab.json"))
res = {}


def dump():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res), flush=True)


def run(ndofs):
    nx, ny, nz, r = pt.host.cube_sizing(ndofs, True, 3, 1, 1)
    f = 2 ** r
    P = pt.host.Problem("elasticity", 1, nx * f, ny * f, nz * f)
    tag = f"{P.n_owned * 3}"
    res[tag] = {"box": [nx * f, ny * f, nz * f], "nnz_blocks": int(P.nnz)}
    ref = None
    only = os.environ.get("RING_AB_ONLY")  # e.g. "ring" under ncu
    for name, env in (("ring", {}), ("walk3", {"PTB_ASM_RING": "0"}),
                      ("cellorder", {"PTB_ASM_RING": "0", "PTB_ASM_WALK3": "0"})):
        if only and name != only:
            continue
        for k in ("PTB_ASM_RING", "PTB_ASM_WALK3", "PTB_RING_WARPS"):
            os.environ.pop(k, None)
        os.environ.update(env)
        t0 = time.perf_counter()
        c = pt.abi.Context(0)
        c.set_problem(P)
        res[tag][f"{name}_set_problem_s"] = time.perf_counter() - t0
        c.assemble_matrix()
        a = c.matrix_values()
        if ref is None:
            ref = a
            rows = np.repeat(np.arange(P.n_owned), np.diff(P["rowptr"]))
            own = P["cols"] == rows
            diag = np.zeros(P.n_owned)
            diag[rows[own]] = np.abs(a.reshape(-1, 9)[own]).max(axis=1)
            scale = diag[rows][:, None]
        else:
            res[tag][f"{name}_vs_ring_max_rel_row_diag"] = float((np.abs(a - ref).reshape(-1, 9) / scale).max())
        if name == "ring":
            for warps in ((4, 2, 1) if not only else ()):
                os.environ["PTB_RING_WARPS"] = str(warps)
                res[tag][f"ring_warps{warps}_ms"] = c.time_kernel(pt.abi.KERNEL_ASSEMBLE_MATRIX, 5)
            os.environ.pop("PTB_RING_WARPS", None)
            if only:
                res[tag]["ring_ms"] = c.time_kernel(pt.abi.KERNEL_ASSEMBLE_MATRIX, 2)
        else:
            res[tag][f"{name}_ms"] = c.time_kernel(pt.abi.KERNEL_ASSEMBLE_MATRIX, 5)
        res[tag][f"{name}_device_bytes"] = int(c.device_bytes())
        dump()
        c.close()


if __name__ == "__main__":
    for n in [int(a) for a in sys.argv[1:]] or [10_000_000]:
        run(n)
