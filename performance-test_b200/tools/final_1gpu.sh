set -u
mkdir -p gpurun_out/f1
(timeout 900 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/f1/gpu_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/f1/gpu_tests.log)
tail -14 gpurun_out/f1/gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
/usr/bin/time -f "bench default wall %e s" timeout 900 python bench.py > gpurun_out/f1/bench_default.json 2> gpurun_out/f1/bench_default.err; tail -2 gpurun_out/f1/bench_default.err
/usr/bin/time -f "bench reference wall %e s" timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f1/bench_reference.json 2> gpurun_out/f1/bench_reference.err; tail -1 gpurun_out/f1/bench_reference.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/f1/bench_default.json"))
def show(d,tag):
    r=d["roofline"]; print(tag, "value %.4g e2e %.4g"%(d["value"],d["e2e"]["value"]), d["stage_ms"], "its", d["cg_iterations"], "spmv ms %.4f frac %.3f"%(r["ms_per_launch"],r["frac"]), r["cg_iteration"], "nnz/s %.4g"%d["assembled_nnz_per_s"], "asm frac %.3f"%r["other_kernels"]["assemble_matrix"]["frac_of_hbm_peak"])
    for k,v in d.get("renumbered",{}).items(): print("   renumbered", k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a!="stage_ms"})
show(d,"HEAD"); show(d["secondary"],"SEC"); print(d["cpu_baseline"])
r=json.load(open("gpurun_out/f1/bench_reference.json")); print("REF value %.4g"%r["value"], r["cpu_baseline"]["cores"], r["cpu_baseline"]["sample"][:260])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/f1/launches_elasticity_10M.csv python bench.py --workload elasticity --steps 1 --warmup 1 --no-cpu-baseline --no-renumbered --kmax 60 > /dev/null 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:spmv_sell -s 30 -c 1 -f -o gpurun_out/f1/ncu_spmv_elasticity_10M python bench.py --workload elasticity --steps 1 --warmup 0 --no-cpu-baseline --no-renumbered --kmax 40 > /dev/null 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:assemble_matrix_p1_walk3 -c 1 -f -o gpurun_out/f1/ncu_walk3_elasticity_10M python bench.py --workload elasticity --steps 1 --warmup 0 --no-cpu-baseline --no-renumbered --kmax 10 > /dev/null 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:cg_loop -c 1 -f -o gpurun_out/f1/ncu_cg_loop_elasticity_1250k python bench.py --workload elasticity --ndofs 1250000 --steps 1 --warmup 0 --no-cpu-baseline --no-renumbered --kmax 100 > /dev/null 2>&1
ls -la gpurun_out/f1/
