"""Generates tests/golden/ref_cg.json: answers of the REFERENCE'S linalg::cg (src/cg.h:38-86,
compiled unchanged into oracle/_ref/libref.so) on assembled operators of the host stand-in:
iteration count, |x|_2, and 16 sampled entries of x. The matrix and right-hand side fed to it are
the oracle's (the reference gets them from DOLFINx/PETSc, which cannot be built here), so these
vectors pin the SOLVE loop (R13-R15), not the assembly. Run here: python tests/golden/make_ref_cg.py
"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CASES = [("poisson P1 12x11x13", "poisson", 1, (12, 11, 13), 5000, 1e-8),
         ("poisson P1 kmax 7", "poisson", 1, (9, 8, 7), 7, 1e-30),
         ("poisson P2 5x4x6", "poisson", 2, (5, 4, 6), 5000, 1e-8),
         ("poisson P3 3x4x3 cgpoisson settings", "poisson", 3, (3, 4, 3), 100, 1e-6),
         ("elasticity P1 7x6x8", "elasticity", 1, (7, 6, 8), 5000, 1e-8)]

if __name__ == "__main__":
    pt = importlib.import_module("performance-test_b200")
    import oracle
    from oracle import ref
    if not ref.build():
        raise SystemExit("needs /root/reference (run in the build container)")
    oracle.build()
    out = []
    for name, ptype, order, dims, kmax, rtol in CASES:
        P = pt.host.Problem(ptype, order, *dims)
        A, b = oracle.assemble_matrix(P), oracle.assemble_vector(P)
        xs, k = ref.cg([dict(bs=P.bs, n_owned=P.n_owned, n_ghost=0, rowptr=P["rowptr"], cols=P["cols"],
                             vals=A, b=b)], kmax=kmax, rtol=rtol)
        x = xs[0]
        idx = np.linspace(0, len(x) - 1, 16).astype(int)
        out.append(dict(name=name, ptype=ptype, order=order, dims=list(dims), kmax=kmax, rtol=rtol,
                        iterations=k, x_norm=float(np.linalg.norm(x)), x_absmax=float(np.abs(x).max()),
                        sample_idx=idx.tolist(), x_sample=[float(v) for v in x[idx]]))
        print(name, "iterations", k)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_cg.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path)
