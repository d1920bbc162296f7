"""Generates tests/golden/sizing.json FROM THE REFERENCE'S OWN CODE: oracle/_ref/libref.so holds
src/mesh.cpp:44-74 (num_entities / num_pdofs) and :82-151 (the create_cube_mesh search) compiled
from /root/reference by oracle/ref/Makefile; this script calls it for the BASELINE.json configs, the
nine CI configurations (.github/workflows/ccpp.yml:56-197) and a few edge cases, and stores the
answers so that the GPU box (no /root/reference) and later rounds check against fixed numbers.
A pure-Python replay of the same arithmetic runs next to it as a second route; the two must agree
before anything is written. Run here: python tests/golden/make_sizing.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def num_entities(i, j, k, nrefine):
    i <<= nrefine; j <<= nrefine; k <<= nrefine
    return ((i + 1) * (j + 1) * (k + 1), 7 * i * j * k + 3 * (i * j + i * k + j * k) + (i + j + k),
            12 * i * j * k + 2 * (i * j + i * k + j * k), 6 * i * j * k)


def num_pdofs(i, j, k, nrefine, order):
    nv, ne, nf, nc = num_entities(i, j, k, nrefine)
    return {1: nv, 2: nv + ne, 3: nv + 2 * ne + nf, 4: nv + 3 * ne + 3 * nf + nc}[order]


def sizing(target, total, dofs_per_node, order, nproc):
    N = target // dofs_per_node if total else target * nproc // dofs_per_node
    Nx, r, ndofs = 1, 0, 0
    while ndofs < N:
        Nx += 1
        if Nx > 200:
            while ndofs < N:
                r += 1
                ndofs = num_pdofs(Nx, Nx, Nx, r, order)
            while ndofs > N:
                Nx -= 1
                ndofs = num_pdofs(Nx, Nx, Nx, r, order)
        ndofs = num_pdofs(Nx, Nx, Nx, r, order)
    Ny = Nz = Nx
    mindiff = 1000000
    for i in range(Nx - 10, Nx + 10):
        for j in range(i - 5, i + 5):
            for k in range(i - 5, i + 5):
                diff = abs(num_pdofs(i, j, k, r, order) - N)
                if diff < mindiff:
                    mindiff, best = diff, (i, j, k)
    if mindiff < 1000000:
        Nx, Ny, Nz = best
    return [Nx, Ny, Nz, r]


CASES = [  # (name, target, total, dofs_per_node, order, nproc)
    ("C1 poisson P1 500k", 500000, False, 1, 1, 1),
    *[(f"C2 poisson P1 weak 20M x{p}", 20000000, False, 1, 1, p) for p in (1, 2, 4, 8)],
    *[(f"C3 elasticity P1 strong 10M x{p}", 10000000, True, 3, 1, p) for p in (1, 2, 4, 8)],
    ("C4 poisson P2 50M", 50000000, True, 1, 2, 8), ("C4 poisson P3 50M", 50000000, True, 1, 3, 8),
    *[(f"C5 elasticity P1 weak 100M x{p}", 100000000, False, 3, 1, p) for p in (1, 2, 4, 8)],
    ("CI poisson serial", 50000, False, 1, 1, 1), ("CI poisson weak np2", 50000, False, 1, 1, 2),
    ("CI poisson P3 weak np2", 50000, False, 1, 3, 2), ("CI poisson strong np2", 1000000, True, 1, 1, 2),
    ("CI elasticity serial", 100000, False, 3, 1, 1), ("CI elasticity weak np2", 100000, False, 3, 1, 2),
    ("CI elasticity P3 weak np2", 100000, False, 3, 3, 2), ("CI elasticity strong np2", 500000, True, 3, 1, 2),
    ("default --ndofs 50000 P2", 50000, False, 1, 2, 1), ("tiny", 10, True, 1, 1, 1),
    ("P4 (counts only in the reference)", 200000, True, 1, 4, 1),
    ("Nx_max boundary P1", 8120601, True, 1, 1, 1), ("just above Nx_max P1", 8300000, True, 1, 1, 1),
    ("P2 refined", 300000000, True, 1, 2, 1), ("weak x64 elasticity P2", 500000, False, 3, 2, 64),
]

if __name__ == "__main__":
    from oracle import ref
    if not ref.build():
        raise SystemExit("needs /root/reference (run in the build container)")
    out = []
    for name, target, total, dpn, order, nproc in CASES:
        got = list(ref.cube_sizing(target, total, dpn, order, nproc))
        assert got == sizing(target, total, dpn, order, nproc), (name, got)
        Nx, Ny, Nz, r = got
        ent = list(ref.num_entities(Nx, Ny, Nz, r))
        assert ent == list(num_entities(Nx, Ny, Nz, r))
        pd = ref.num_pdofs(Nx, Ny, Nz, r, order)
        assert pd == num_pdofs(Nx, Ny, Nz, r, order)
        out.append(dict(name=name, target=target, total=total, dofs_per_node=dpn, order=order,
                        nproc=nproc, sizing=got, entities=ent, pdofs=pd))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sizing.json")
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, len(out), "cases (source: oracle/_ref/libref.so = src/mesh.cpp compiled here)")
