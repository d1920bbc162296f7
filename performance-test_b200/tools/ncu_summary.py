"""Summarise an .ncu-rep (read locally, no GPU needed) into the CSV layout kept under profiles/.

    python performance-test_b200/tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r02_x.csv [kernel-regex]
    python performance-test_b200/tools/ncu_summary.py gpurun_out/prof.ncu-rep --source kernel-regex

The second form prints, for the first matching kernel, the SASS regions that share an execution
count (prologue / unrolled loop bodies / epilogue) with their share of the stall samples and the
three leading stall reasons -- the view that located the prologue problem of the walk kernel.
"""
import csv
import io
import re
import subprocess
import sys

METRICS = [
    "Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__warps_active.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True,
                         text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def summary(rep, dst, pattern):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    keep = [m for m in METRICS if m in hdr]
    idx = [hdr.index(m) for m in keep]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(keep)
        w.writerow([units[i] for i in idx])
        for r in rows[2:]:
            if len(r) > max(idx) and re.search(pattern, r[hdr.index("Kernel Name")]):
                w.writerow([r[i] for i in idx])
    print("wrote", dst)


def source(rep, pattern):
    rows = ncu_csv(rep, "source", ["--kernel-name", f"regex:{pattern}"])
    hdr = rows[1]
    isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
    stalls = {h: hdr.index(h) for h in hdr if h.startswith("stall_") and "Not Issued" not in h}
    R = [r for r in rows[2:] if len(r) > isamp and r[isamp].isdigit()]
    total = sum(int(r[isamp]) for r in R) or 1
    seg = None
    print(rows[0][1] if len(rows[0]) > 1 else "", "-", len(R), "instructions,", total, "samples")
    for i, r in enumerate(R + [None]):
        ex = int(r[iex]) if r else -1
        if seg is None or r is None or abs(ex - seg["ex"]) > 0.03 * max(ex, seg["ex"], 1):
            if seg and (seg["s"] > 0.004 * total or seg["n"] > 15):
                top = sorted(seg["st"].items(), key=lambda kv: -kv[1])[:3]
                print(f"i={seg['i']:5d} n={seg['n']:4d} exec/inst={seg['ex']:>9d} "
                      f"samples={100 * seg['s'] / total:5.1f}%  "
                      + ", ".join(f"{k[6:]} {100 * v // max(seg['s'], 1)}%" for k, v in top))
            if r is None:
                break
            seg = {"i": i, "ex": ex, "n": 0, "s": 0, "st": {h: 0 for h in stalls}}
        seg["n"] += 1
        seg["s"] += int(r[isamp])
        for h, j in stalls.items():
            seg["st"][h] += int(r[j])


if __name__ == "__main__":
    if len(sys.argv) >= 4 and sys.argv[2] == "--source":
        source(sys.argv[1], sys.argv[3])
    elif len(sys.argv) >= 3:
        summary(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ".")
    else:
        print(__doc__)
