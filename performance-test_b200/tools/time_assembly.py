"""Kernel times of the assembly stages on resident data (ptb_time_kernel, CUDA events), one line of JSON.
    python performance-test_b200/tools/time_assembly.py <poisson|elasticity> <order> <ndofs> [reps]"""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
pt = importlib.import_module("performance-test_b200")

ptype, order, ndofs = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
bs = 3 if ptype == "elasticity" else 1
nx, ny, nz, r = pt.host.cube_sizing(ndofs, True, bs, order, 1)
f = 2 ** r
P = pt.host.Problem(ptype, order, nx * f, ny * f, nz * f)
c = pt.abi.Context(0)
c.set_problem(P)
c.assemble_matrix()
c.assemble_vector()
out = {"ptype": ptype, "order": order, "dofs": P.n_owned * bs, "nnz": int(P.nnz) * bs * bs,
       "matrix_ms": c.time_kernel(pt.abi.KERNEL_ASSEMBLE_MATRIX, reps),
       "vector_ms": c.time_kernel(pt.abi.KERNEL_ASSEMBLE_VECTOR, reps),
       "env": {k: v for k, v in os.environ.items() if k.startswith("PTB_")}}
out["nnz_per_s"] = out["nnz"] / (out["matrix_ms"] * 1e-3)
print(json.dumps(out))
c.close()
