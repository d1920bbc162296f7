"""dolfinx-scaling-test: the reference's command-line surface (src/main.cpp:54-115,170,186-205,
226,232-233). CPU part: option handling and error exits; GPU part: a full run whose iteration
count must match the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "performance-test_b200", "dolfinx-scaling-test")


def _run(*args, timeout=300):
    return subprocess.run([EXE, *args], capture_output=True, text=True, timeout=timeout)


def test_help_lists_the_reference_options(pt):
    r = _run("--help")
    assert r.returncode == 0
    for opt in ("--problem_type", "--mesh_type", "--memory_profiling", "--subcomm_partition",
                "--scaling_type", "--output", "--ndofs", "--order", "--scatterer"):
        assert opt in r.stdout
    assert "(=poisson)" in r.stdout and "(=50000)" in r.stdout and "(=weak)" in r.stdout


@pytest.mark.parametrize("args,msg", [
    (["--scaling_type", "sideways"], "Scaling type 'sideways` unknown"),   # main.cpp:115
    (["--problem_type", "stokes"], "Unknown problem type: stokes"),        # main.cpp:170
    (["-pc_type", "gamg"], "AMG is out of scope"),
    (["--ndofs"], "required argument"),
])
def test_bad_options_exit_non_zero(pt, args, msg):
    r = _run(*args)
    assert r.returncode != 0
    assert msg in r.stderr


def test_unregistered_options_are_ignored_until_the_gpu_is_needed(pt):
    """allow_unregistered (main.cpp:76-82): unknown options do not fail parsing. Without a GPU the
    run then stops at ptb_create with the no-fallback message; with one it completes."""
    r = _run("--ndofs", "2000", "-ksp_view", "-log_view", "--some_future_flag=3")
    assert "UnitCube (10x12x13) to be refined 0 times" in r.stdout  # mesh.cpp:190-194 + sizing
    if r.returncode != 0:
        assert "no CUDA device (there is no CPU fallback)" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("ptype,ndofs", [("poisson", 40000), ("elasticity", 30000)])
def test_full_run_matches_oracle(pt, oracle, ptype, ndofs):
    r = _run("--problem_type", ptype, "--ndofs", str(ndofs), "-ksp_rtol", "1e-8")
    assert r.returncode == 0, r.stderr
    out = r.stdout
    for row in ("ZZZ Create Mesh", "ZZZ FunctionSpace", "ZZZ Create boundary conditions",
                "ZZZ Create RHS function", "ZZZ Assemble matrix", "ZZZ Assemble vector", "ZZZ Solve"):
        assert row in out
    # the enclosing "ZZZ Assemble" timer exists for poisson only (poisson_problem.cpp:49,159)
    assert bool(re.search(r"^ZZZ Assemble\s+\|", out, re.M)) == (ptype == "poisson")
    assert "Test problem summary" in out and f"Problem type:    {ptype}" in out
    its = int(re.search(r"\*\*\* Number of Krylov iterations: (\d+)", out).group(1))
    norm = float(re.search(r"\*\*\* Solution norm:\s+([0-9.eE+-]+)", out).group(1))
    dpn = 3 if ptype == "elasticity" else 1
    Nx, Ny, Nz, rr = pt.host.cube_sizing(ndofs, False, dpn, 1, 1)
    assert f"UnitCube ({Nx}x{Ny}x{Nz}) to be refined {rr} times" in out
    P = pt.host.Problem(ptype, 1, Nx, Ny, Nz)
    assert f"Total degrees of freedom:               {P.n_global * P.bs}" in out
    A, b = oracle.assemble_matrix(P), oracle.assemble_vector(P)
    x, k, _ = oracle.cg(P.bs, P.n_owned, P["rowptr"], P["cols"], A, b, kmax=10000, rtol=1e-8,
                        precond="jacobi")
    assert abs(its - k) <= 1
    assert norm == pytest.approx(np.linalg.norm(x), rel=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("ptype,order,ndofs", [("poisson", 1, 40000), ("elasticity", 1, 30000), ("poisson", 2, 30000)])
def test_device_setup_run_equals_the_host_setup_run(pt, ptype, order, ndofs):
    """--device_setup (mesh, dofmap, pattern, boundary conditions, RHS generated on the GPU; with
    PTB_GPU_SETUP=1 the layouts too) must print the same iteration count and solution norm as the
    default run whose setup arrays come from the host stand-in."""
    out = {}
    for flag in ([], ["--device_setup"]):
        env = dict(os.environ, PTB_GPU_SETUP="1" if flag else "0")
        r = subprocess.run([EXE, "--problem_type", ptype, "--order", str(order), "--ndofs", str(ndofs),
                            "-ksp_rtol", "1e-8", *flag], capture_output=True, text=True, timeout=300, env=env)
        assert r.returncode == 0, r.stderr
        its = int(re.search(r"\*\*\* Number of Krylov iterations: (\d+)", r.stdout).group(1))
        norm = float(re.search(r"\*\*\* Solution norm:\s+([0-9.eE+-]+)", r.stdout).group(1))
        out[bool(flag)] = (its, norm)
    assert abs(out[True][0] - out[False][0]) <= 1
    assert out[True][1] == pytest.approx(out[False][1], rel=1e-6)


@pytest.mark.gpu
def test_cgpoisson_uses_cg_h_defaults(pt):
    """cgpoisson: linalg::cg(u, b, action, 100, 1e-6), no preconditioner (cgpoisson_problem.cpp:233)."""
    r = _run("--problem_type", "cgpoisson", "--ndofs", "200000")
    assert r.returncode == 0, r.stderr
    its = int(re.search(r"\*\*\* Number of Krylov iterations: (\d+)", r.stdout).group(1))
    assert 1 <= its <= 100
    assert "Gdof/s" in r.stdout


@pytest.mark.gpu
def test_reference_ci_configuration_elasticity_order_3(pt):
    """The reference's CI runs `--problem_type elasticity --scaling_type weak --ndofs 100000 --order 3`
    (.github/workflows/ccpp.yml:165-181, with GAMG); the same command line with CG + Jacobi must run
    through, print the ZZZ rows and a converged iteration count."""
    r = subprocess.run([EXE, "--problem_type", "elasticity", "--scaling_type", "weak", "--ndofs", "100000",
                        "--order", "3", "-ksp_rtol", "1e-8", "-pc_type", "jacobi"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ZZZ Assemble matrix" in r.stdout and "ZZZ Solve" in r.stdout
    its = int(re.search(r"\*\*\* Number of Krylov iterations: (\d+)", r.stdout).group(1))
    assert 10 < its < 10000

