"""The generated reference tensors the CUDA P2/P3 kernels use (csrc/element_tables.h, exact
monomial integration in 60-digit arithmetic) against the oracle's route (nodal basis tabulated at
Gauss-Jacobi points): two independent derivations of the same FFCx/Basix tensors. CPU only."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from test_oracle_kats import _single_tet

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "performance-test_b200", "csrc", "element_tables.h")


def _arr(name):
    src = open(HDR).read()
    m = re.search(r"static const double %s\[\d+\] = \{(.*?)\};" % name, src, re.S)
    return np.array([float(x) for x in m.group(1).replace("\n", " ").split(",") if x.strip()])


def test_header_is_what_the_generator_writes(tmp_path):
    out = tmp_path / "element_tables.h"
    subprocess.run([sys.executable, os.path.join(ROOT, "performance-test_b200", "tools",
                                                 "gen_element_tables.py"), str(out)], check=True,
                   capture_output=True)
    assert out.read_text() == open(HDR).read()


@pytest.mark.parametrize("order,nd", [(2, 10), (3, 20)])
def test_tensors_match_oracle_quadrature(oracle, order, nd):
    rng = np.random.default_rng(11)
    X = np.array([[0., 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]]) + 0.2 * rng.standard_normal((4, 3))
    S, M = _arr(f"S_P{order}").reshape(6, nd, nd), _arr(f"M_P{order}").reshape(nd, nd)
    MF = _arr(f"MF_P{order}").reshape(4, nd, nd)
    P, A = _single_tet(oracle, "poisson", order, X)
    e = X[1:] - X[0]
    c = np.array([np.cross(e[1], e[2]), np.cross(e[2], e[0]), np.cross(e[0], e[1])])
    det = e[0] @ c[0]
    G = c @ c.T / abs(det)
    Ae = (G[0, 0] * S[0] + G[0, 1] * S[1] + G[0, 2] * S[2] + G[1, 1] * S[3] + G[1, 2] * S[4]
          + G[2, 2] * S[5])
    Ae_ref = oracle.assemble_matrix(P).reshape(nd, nd)
    assert np.abs(Ae - Ae_ref).max() <= 1e-13 * np.abs(Ae_ref).max()
    f, g = rng.standard_normal(nd), rng.standard_normal(nd)
    A["f"], A["g"] = f, np.zeros(0)
    be_ref = oracle.assemble_vector(P)
    assert np.abs(abs(det) * M @ f - be_ref).max() <= 1e-13 * np.abs(be_ref).max()
    # facets: every face of a single tet is exterior
    A["f"], A["g"] = np.zeros(nd), g
    A["facet_cells"], A["facet_local"] = np.zeros(4, np.int32), np.arange(4, dtype=np.int32)
    bf_ref = oracle.assemble_vector(P)
    faces = [(1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2)]
    bf = np.zeros(nd)
    for lf, (a, b, cc) in enumerate(faces):
        scale = np.linalg.norm(np.cross(X[b] - X[a], X[cc] - X[a]))
        bf += scale * MF[lf] @ g
    assert np.abs(bf - bf_ref).max() <= 1e-13 * np.abs(bf_ref).max()
