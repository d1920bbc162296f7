set -u
mkdir -p gpurun_out/f2
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "balanced or persistent or cg_matches" 2>&1 | tail -2
W="--workload elasticity --ndofs 1250000 --steps 2 --warmup 1"; performance-test_b200/tools/ab_cg.sh gpurun_out/f2/e1250k $W -- "resident:" "res0:PTB_LOOP_RESIDENT=0"
PTB_LOOP_TRACE=40 python bench.py --workload elasticity --ndofs 1250000 --steps 1 --warmup 0 --no-cpu-baseline --no-renumbered 2>&1 >/dev/null | grep "loop trace" | tail -8
t0=$(date +%s); timeout 900 python bench.py > gpurun_out/f2/bench_default.json 2> gpurun_out/f2/bench_default.err; echo "bench default wall $(( $(date +%s) - t0 )) s"; tail -2 gpurun_out/f2/bench_default.err
t0=$(date +%s); timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f2/bench_reference.json 2> gpurun_out/f2/bench_reference.err; echo "bench reference wall $(( $(date +%s) - t0 )) s"; tail -1 gpurun_out/f2/bench_reference.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/f2/bench_default.json"))
def show(d,tag):
    r=d["roofline"]; print(tag, "value %.4g e2e %.4g"%(d["value"],d["e2e"]["value"]), d["stage_ms"], "its", d["cg_iterations"], "spmv ms %.4f frac %.3f"%(r["ms_per_launch"],r["frac"]), r["cg_iteration"], "nnz/s %.4g"%d["assembled_nnz_per_s"], "asm frac %.3f"%r["other_kernels"]["assemble_matrix"]["frac_of_hbm_peak"])
    for k,v in d.get("renumbered",{}).items(): print("   renumbered", k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a!="stage_ms"})
show(d,"HEAD"); show(d["secondary"],"SEC"); print(d["cpu_baseline"])
r=json.load(open("gpurun_out/f2/bench_reference.json")); print("REF value %.4g"%r["value"], r["cpu_baseline"]["cores"], r["cpu_baseline"]["sample"][:300])
PY
