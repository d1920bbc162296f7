#!/usr/bin/env bash
# First GPU job after a round that ended with untested opt-in kernels:
# parity first, then A/B timings. Everything lands in gpurun_out/first_call/. One GPU, ~12 minutes
# (the parity tests are the first two; every later step has its own timeout and can be cut).
#   gpurun --timeout 1200 -- 'bash performance-test_b200/tools/first_gpu_call.sh'
set -u
cd "$(dirname "$0")/../.."
out=gpurun_out/first_call
mkdir -p "$out"
export PTB_TEST_OPTIN=1
echo "== opt-in parity tests"
timeout 400 python -m pytest tests/test_gpu_parity.py -q -s -k "opt_in or persistent or star_walk or binned or compaction" 2>&1 | tail -40 | tee "$out/optin_tests.txt"
timeout 200 python -m pytest tests/test_cli.py tests/test_gpu_multi.py -q -k "device_setup or device_generated" 2>&1 | tail -5 | tee -a "$out/optin_tests.txt"
echo "== the reference's timing table with the setup on the host and on the device (Poisson 4M DOFs)"
for ds in "" "--device_setup"; do
  PTB_GPU_SETUP=1 timeout 200 performance-test_b200/dolfinx-scaling-test --ndofs 4000000 -ksp_rtol 1e-8 $ds \
    > "$out/cli_poisson4M${ds:+_device_setup}.txt" 2>&1
  grep -E "ZZZ|Krylov|Solution norm" "$out/cli_poisson4M${ds:+_device_setup}.txt" | head -20
done
echo "== assembly A/B (4M DOFs)"
WALK_CHECK_OUT=first_call/assembly_ab_4M.json timeout 200 python performance-test_b200/tools/check_walk.py ab2 4000000 2>&1 | tail -2
echo "== ncu --set full of the direct-gather kernels (Poisson 4M): read it with tools/ncu_summary.py"
PTB_ASM_GWALK=1 timeout 200 ncu --set full --import-source on --clock-control none -k regex:gwalk -c 2 -f \
  -o "$out/prof_gwalk" python performance-test_b200/tools/check_walk.py ncu2 4000000 2>&1 | tail -3
echo "== matrix-free CG (cgpoisson action), staged star vs direct-gather walk (4M DOFs)"
WALK_CHECK_OUT=first_call/matrix_free_4M.json timeout 200 python performance-test_b200/tools/check_walk.py abmf 4000000 2>&1 | tail -2
echo "== P2/P3 matrix assembly: all slices vs row-length bins (2M DOFs)"
WALK_CHECK_OUT=first_call/assembly_pk_2M.json timeout 200 python performance-test_b200/tools/check_walk.py abpk 2000000 2>&1 | tail -2
echo "== P1 assembly maps built on the host vs on the device (PTB_GPU_SETUP, 4M DOFs)"
WALK_CHECK_OUT=first_call/setup_4M.json timeout 200 python performance-test_b200/tools/check_walk.py absetup 4000000 2>&1 | tail -2
echo "== bench line with the whole P1 setup generated on the device (20M DOFs): setup_s and value against the default line"
PTB_GPU_SETUP=1 PTB_BENCH_DEVICE_SETUP=1 timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline \
  > "$out/bench_poisson20M_device_setup.json" 2> "$out/bench_poisson20M_device_setup.err"
python -c "import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], 'value %.4g' % d['value'], d['setup_s'], d['stage_ms'], 'its', d['cg_iterations'])" "$out/bench_poisson20M_device_setup.json" 2>/dev/null || echo "FAILED (see $out/bench_poisson20M_device_setup.err)"
echo "== SpMV on the zero-compacted operator (Poisson 20M: the headline line), off / on"
for z in "0 0" "1 0" "1 1e-14"; do
  set -- $z
  PTB_SPMV_COMPACT=$1 PTB_SPMV_COMPACT_TOL=$2 timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu-baseline \
    > "$out/bench_poisson20M_compact$1_tol$2.json" 2> "$out/bench_poisson20M_compact$1_tol$2.err"
  python -c "import json,sys; d=json.load(open(sys.argv[1])); r=d['roofline']; print(sys.argv[1], 'value %.4g' % d['value'], d['stage_ms'], 'spmv ms', r['ms_per_launch'], 'stored/pattern', r['spmv_stored_entries'], r['pattern_entries'])" "$out/bench_poisson20M_compact$1_tol$2.json" 2>/dev/null || echo "FAILED (see the .err file)"
done
echo "== CG loop: three kernels per iteration vs persistent kernel, small problem (config 1) and 3M DOFs"
for persistent in 0 1; do
  for wl in "--workload small" "--workload poisson --ndofs 3000000" "--workload elasticity --ndofs 1250000"; do
    tag=$(echo "$wl" | tr -d ' -' | tr -c 'a-z0-9\n' '_')
    PTB_CG_PERSISTENT=$persistent timeout 200 python bench.py $wl --steps 2 --warmup 1 --no-cpu-baseline \
      > "$out/bench_${tag}_persistent${persistent}.json" 2> "$out/bench_${tag}_persistent${persistent}.err"
    python - "$out/bench_${tag}_persistent${persistent}.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], "value %.4g" % d["value"], "its", d.get("cg_iterations"), "launches", d.get("gpu_launches"), d.get("stage_ms"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  done
done

# Two GPUs (gpurun --gpus 2): the existing 2-rank tests with the persistent loop switched on; the
# spawned ranks inherit the environment, so no extra test code is needed:
#   PTB_CG_PERSISTENT=1 timeout 300 python -m pytest tests/test_gpu_multi.py -q -k peer
#   PTB_TEST_DEVICE_SETUP=1 PTB_GPU_SETUP=1 timeout 300 python -m pytest tests/test_gpu_multi.py -q -k two_rank
# then the strong-scaling point that motivated the kernel:
#   for p in 0 1; do PTB_CG_PERSISTENT=$p python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
#     --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --workload elasticity --steps 2 --warmup 1; done
