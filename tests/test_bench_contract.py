"""bench.py contract pieces that can be checked without a GPU: the reference arm prints one JSON
line with the agreed keys, and the algorithmic byte model matches SURVEY 8(d)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--steps", "1", "--warmup", "0", "--ndofs", "20000", "--workload", "elasticity",
                        "--cpu-kcap", "40"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "cg_dof_iters_per_s"
    assert j["unit"] == "DOF-iters/s" and j["higher_is_better"] is True and j["dtype"] == "f64"
    assert j["value"] > 0 and j["cg_iterations"] == 40
    assert "same mesh as the GPU arm" in j["cpu_baseline"]["sample"]
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"] == {"value": j["value"], "unit": "DOF-iters/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert "workload" in j["config"]


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--gpus", "2"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_algorithmic_byte_model(pt):
    sys.path.insert(0, ROOT)
    import bench
    P = pt.host.Problem("poisson", 1, 6, 5, 4)
    spmv, cg, asm = bench.algorithmic_bytes(P)
    assert spmv == 12 * P.nnz + 20 * P.n_owned            # scalar CSR
    assert cg == spmv + 96 * P.n_owned                     # + 96 B/DOF of vector traffic (Jacobi)
    E = pt.host.Problem("elasticity", 1, 4, 3, 3)
    spmv, cg, asm = bench.algorithmic_bytes(E)
    assert spmv == 76 * E.nnz + 52 * E.n_owned             # 3x3 BCSR
    assert cg == spmv + 96 * 3 * E.n_owned


def test_default_run_is_the_north_star_target():
    """No flags = BASELINE configs[2] (elasticity P1 strong 10M) as the headline line, configs[1]
    (Poisson P1 weak 20M/GPU) under "secondary"."""
    sys.path.insert(0, ROOT)
    import bench
    assert bench.DEFAULT_HEADLINE == "elasticity" and bench.DEFAULT_SECONDARY == "poisson"
    assert bench.WORKLOADS["elasticity"][:4] == ("elasticity", "strong", 10_000_000, 1)
    assert bench.WORKLOADS["poisson"][:4] == ("poisson", "weak", 20_000_000, 1)
    assert bench.WORKLOADS["elasticity_weak"][:4] == ("elasticity", "weak", 100_000_000, 1)
