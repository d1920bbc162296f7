#!/usr/bin/env bash
# A/B of the solve-stage switches on one workload: every variant is one bench.py process (the
# switches are read once per process). Prints one line per variant and keeps the JSON lines.
#   tools/ab_cg.sh <out_dir> <bench args...> -- "NAME:ENV=V ENV2=V" "NAME2:" ...
# With N > 1 GPUs: set NGPU=N (torch.distributed.run is used).
set -u
cd "$(dirname "$0")/../.."
out=$1; shift
args=()
while [ $# -gt 0 ] && [ "$1" != "--" ]; do args+=("$1"); shift; done
shift
mkdir -p "$out"
n=${NGPU:-1}
for v in "$@"; do
  name=${v%%:*}; envs=${v#*:}
  if [ "$n" -gt 1 ]; then
    launcher=(python -m torch.distributed.run --nnodes=1 --nproc-per-node "$n" --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus "$n")
  else
    launcher=(python bench.py)
  fi
  env $envs timeout 600 "${launcher[@]}" "${args[@]}" --no-cpu-baseline --no-renumbered > "$out/$name.json" 2> "$out/$name.err"
  python - "$out/$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    r = d["roofline"]; it = r["cg_iteration"]
    print("%-28s value %.4g  its %d  it_us %.1f  spmv_us %.1f (frac %.3f)  upd_us %.1f  dir_us %.1f  launches %d  asm_ms %.3f" % (
        sys.argv[2], d["value"], d["cg_iterations"], 1e3 * it["ms_measured"], 1e3 * r["ms_per_launch"], r["frac"],
        1e3 * r["other_kernels"]["cg_update"]["ms"], 1e3 * r["other_kernels"]["cg_direction"]["ms"],
        d["gpu_launches"], d["stage_ms"]["assemble_matrix"]))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done
