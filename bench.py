#!/usr/bin/env python
"""bench.py -- the hot path of FEniCS/performance-test on B200: ZZZ Assemble matrix +
ZZZ Assemble vector + ZZZ Solve (cg.h CG + Jacobi, rtol 1e-8) on the unit-cube tet mesh.

    python bench.py --gpus N --steps K --warmup W                 (N = 1)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        CPU restatement of the reference on the host cores

One "step" = one pass of the hot path on device-resident mesh/dofmap/pattern data: assemble A,
assemble b, solve to rtol 1e-8. Prints ONE JSON line (rank 0):

  metric/value   CG DOF-iterations/s of the ZZZ Solve stage, whole job (the reference's own formula,
                 src/cgpoisson_problem.cpp:236-241: num_it * ndofs_global / t_solve)
  assembled_nnz_per_s   stored CSR nnz / ZZZ Assemble matrix time (second half of BASELINE metric)
  ms_per_step    whole step (assemble matrix + vector + solve), max over ranks
  e2e            same metric through the C-ABI with HOST buffers: per step, coordinates and source
                 terms are copied host->device from pinned memory and b, u are copied back
  roofline       dominant kernel (SpMV inside CG): algorithmic bytes / measured launch time vs the
                 measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline   the CPU oracle ("port": restatement, not DOLFINx/PETSc) on a bounded sample

Workloads (BASELINE.json configs). Default = the north-star target, configs[2] "Elasticity P1
strong scaling 10M DOFs total" as the headline line, and configs[1] "Poisson P1 weak scaling 20M
DOFs/GPU" measured in the same run and reported under "secondary" (same keys). --workload X runs
one workload only: elasticity | poisson | small (configs[0]) | poisson_p2 | poisson_p3 (configs[3])
| elasticity_weak (configs[4], 100M DOFs/GPU, set up on the device).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# torchrun exports OMP_NUM_THREADS=1; the (untimed) host setup is OpenMP code, so give each rank
# its share of the host cores before any OpenMP runtime starts.
_world = int(os.environ.get("WORLD_SIZE", "1"))
if os.environ.get("OMP_NUM_THREADS", "1") == "1":
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // _world))

WORKLOADS = {
    # name: (problem_type, scaling, ndofs, order, description)
    "poisson": ("poisson", "weak", 20_000_000, 1, "Poisson P1 weak scaling 20M DOFs/GPU, CG+Jacobi rtol 1e-8"),
    "elasticity": ("elasticity", "strong", 10_000_000, 1, "Elasticity P1 strong scaling 10M DOFs total, CG+Jacobi rtol 1e-8"),
    "small": ("poisson", "weak", 500_000, 1, "Poisson P1 unit cube 500k DOFs/GPU, CG+Jacobi rtol 1e-8"),
    "poisson_p2": ("poisson", "strong", 50_000_000, 2, "Poisson P2 50M DOFs total (strong), CG+Jacobi rtol 1e-8"),
    "poisson_p3": ("poisson", "strong", 50_000_000, 3, "Poisson P3 50M DOFs total (strong), CG+Jacobi rtol 1e-8"),
    "elasticity_weak": ("elasticity", "weak", 100_000_000, 1, "Elasticity P1 weak scaling 100M DOFs/GPU, CG+Jacobi rtol 1e-8"),
}
DEFAULT_HEADLINE, DEFAULT_SECONDARY = "elasticity", "poisson"
KMAX = 10000  # PETSc's default -ksp_max_it; cg.h's own default of 50 never converges at these sizes


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="one workload only (default: elasticity headline + poisson secondary)")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-renumbered", action="store_true",
                    help="skip the rcm / random dof-numbering measurements of the N = 1 run")
    ap.add_argument("--kmax", type=int, default=KMAX)
    ap.add_argument("--ndofs", type=int, default=None, help="override the workload's --ndofs")
    ap.add_argument("--cpu-kcap", type=int, default=200,
                    help="CG iteration cap of the CPU arm (the metric is a rate)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--comm", default="peer", choices=["peer", "nccl"],
                    help="N>1: NVLink peer-memory kernels (default) or NCCL send/recv + all-reduce")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200", "-i", str(self.device)], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])), mx.append(float(r[2]))
            except Exception:
                continue
            for n, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def sizing(pt, wl, nranks, ndofs_override=None):
    ptype, scaling, ndofs, order, desc = WORKLOADS[wl]
    if ndofs_override:
        ndofs = ndofs_override
    dpn = 3 if ptype == "elasticity" else 1
    Nx, Ny, Nz, r = pt.host.cube_sizing(ndofs, scaling == "strong", dpn, order, nranks)
    return ptype, order, (Nx << r, Ny << r, Nz << r), (Nx, Ny, Nz, r), scaling, ndofs


def algorithmic_bytes(P):
    """SpMV 12*nnz + 20*n (scalar CSR) / 76*nnzb + 52*nb (3x3 BCSR); CG vectors
    96 B/DOF (Jacobi); assembly: values written once + dofmaps + coordinates + markers."""
    n, nnz, bs = P.n_owned, P.nnz, P.bs
    spmv = 12 * nnz + 20 * n if bs == 1 else 76 * nnz + 52 * n
    cg_iter = spmv + 96 * n * bs
    asm = 8 * nnz * bs * bs + P.n_cells * (16 + 4 * P.nd) + 24 * (n + P.n_ghost) + (n + P.n_ghost)
    return spmv, cg_iter, asm


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_arm(pt, wl_name, n_gpus, steps, warmup, kcap, ndofs_override=None):
    """The reference's CPU path for this workload on the host cores, as BASELINE.md section 3 (ii)
    plans it: the SAME mesh and --ndofs as the GPU arm at n_gpus GPUs, split into P partitions on
    P threads -- each thread assembles its rank's rows with the sequential cell loop and the solve
    runs with a halo exchange and rank-ordered reductions (the analogue of `mpirun -np P`). The
    code is the oracle (kind "port": DOLFINx/PETSc cannot be installed here; its CG loop is pinned
    bit for bit to the reference's cg.h compiled into oracle/_ref, tests/test_ref_pin.py), built
    -Ofast like src/CMakeLists.txt:19-20. The solve is capped at kcap iterations: the metric is a
    rate (DOF-iterations/s), and the full solve would take minutes per step."""
    import oracle
    oracle.build()
    ptype, order, dims, base, scaling, ndofs = sizing(pt, wl_name, n_gpus, ndofs_override)
    nthreads = host_threads()
    nparts = max(1, min(nthreads, dims[2]))
    t0 = time.perf_counter()
    probs = [pt.host.Problem(ptype, order, *dims, q, nparts) for q in range(nparts)]
    t_setup = time.perf_counter() - t0
    ndof = probs[0].n_global * probs[0].bs
    nnz = sum(P.nnz * P.bs * P.bs for P in probs)

    def step():
        mats, rhs, t_am, t_av = oracle.assemble_partitions(probs, fast=True)
        t0 = time.perf_counter()
        xs, k, rel = oracle.cg_partitioned(probs, mats, rhs, kmax=kcap, rtol=1e-8, precond="jacobi",
                                           fast=True)
        return t_am, t_av, time.perf_counter() - t0, k

    for _ in range(warmup):
        step()
    ts = [step() for _ in range(steps)]
    t_am, t_av = sum(t[0] for t in ts), sum(t[1] for t in ts)
    t_solve, iters = sum(t[2] for t in ts), sum(t[3] for t in ts)
    wl = WORKLOADS[wl_name]
    sample = (f"same mesh as the GPU arm at {n_gpus} GPU(s): {ptype} P{order} --ndofs {ndofs} "
              f"({scaling}) = {dims[0]}x{dims[1]}x{dims[2]} box, {ndof} DOFs; {nparts} partitions on "
              f"{nparts} threads (z-slabs, halo exchange per iteration); per step: matrix assembly "
              f"{t_am / steps:.2f} s, vector {t_av / steps:.2f} s, first {iters // steps} CG+Jacobi "
              f"iterations (capped at {kcap}; the metric is a rate) {t_solve / steps:.2f} s")
    return {"value": iters * ndof / t_solve, "unit": "DOF-iters/s", "cores": nparts, "kind": "port",
            "sample": sample, "assembled_nnz_per_s": nnz * steps / t_am,
            "ms_per_step": 1e3 * (t_am + t_av + t_solve) / steps, "cg_iterations": iters // steps,
            "same_mesh_as_gpu_arm": True, "ndofs_global": ndof, "host_setup_s": t_setup,
            "workload": wl[4]}


def run_reference(args):
    """--impl reference: the reference's CPU path (cpu_arm above) with all host threads, on the GPU
    arm's own config. Rank 0 alone runs; the other ranks of a torchrun launch exit at once."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pt = importlib.import_module("performance-test_b200")
    wl_name = args.workload or DEFAULT_HEADLINE
    wl = WORKLOADS[wl_name]
    # bounded: the whole --steps K run must end within minutes, so the iteration cap shrinks with K
    kcap = max(10, min(args.cpu_kcap, 1000 // max(args.steps, 1)))
    c = cpu_arm(pt, wl_name, args.gpus, args.steps, min(args.warmup, 1), kcap, args.ndofs)
    line = {
        "impl": "reference", "metric": "cg_dof_iters_per_s", "value": c["value"], "unit": "DOF-iters/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": c["ms_per_step"], "higher_is_better": True,
        "scaling": wl[1], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "assembled_nnz_per_s": c["assembled_nnz_per_s"], "cg_iterations": c["cg_iterations"],
        "ndofs_global": c["ndofs_global"],
        "config": {"workload": wl[4], "cpu_sample": c["sample"]},
        "cpu_baseline": {"value": c["value"], "unit": "DOF-iters/s", "cores": c["cores"],
                         "kind": "port", "sample": c["sample"]},
        "e2e": {"value": c["value"], "unit": "DOF-iters/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    if args.workload is None and not args.no_secondary and args.gpus == 1:
        s = cpu_arm(pt, DEFAULT_SECONDARY, 1, 1, 0, min(kcap, 100))
        line["secondary"] = {"value": s["value"], "unit": "DOF-iters/s", "scaling": "weak",
                             "assembled_nnz_per_s": s["assembled_nnz_per_s"],
                             "cpu_baseline": {"value": s["value"], "unit": "DOF-iters/s",
                                              "cores": s["cores"], "kind": "port", "sample": s["sample"]},
                             "config": {"workload": s["workload"]}}
    print(json.dumps(line), flush=True)


def cpu_baseline(pt, args, wl_name):
    """The in-run baseline of the default line (rank 0, N = 1): one bounded step of cpu_arm."""
    c = cpu_arm(pt, wl_name, 1, 1, 0, min(args.cpu_kcap, 100), args.ndofs)
    return {k: c[k] for k in ("value", "unit", "cores", "kind", "sample", "assembled_nnz_per_s",
                              "same_mesh_as_gpu_arm")}


def measure(pt, env, wl_name, args, steps, warmup, with_cpu, secondary=False):
    """One workload through the C ABI on this rank's GPU; returns the JSON fields (rank 0) or None."""
    import torch
    abi = pt.abi
    world, rank, local_rank, dist = env["world"], env["rank"], env["local_rank"], env["dist"]
    barrier, allmax, allsum = env["barrier"], env["allmax"], env["allsum"]

    # ---- setup (untimed): host mesh/dofmap/pattern, upload, NCCL / peer bootstrap ------------
    ptype, order, dims, base, scaling, ndofs_arg = sizing(pt, wl_name, world, args.ndofs)
    # configs[4] (100M DOFs/GPU) is generated on the device: the host stand-in would need ~10 GB per
    # rank; PTB_BENCH_DEVICE_SETUP=1 forces the same route for any workload
    device_setup = os.environ.get("PTB_BENCH_DEVICE_SETUP") == "1" or wl_name == "elasticity_weak"
    if device_setup:
        os.environ.setdefault("PTB_GPU_SETUP", "1")  # layouts and assembly maps on the device too
    t_setup0 = time.perf_counter()
    P = pt.host.Problem(ptype, order, *dims, rank, world, with_dofmap=not device_setup)
    t_host = time.perf_counter() - t_setup0
    stream = torch.cuda.current_stream().cuda_stream
    ctx = abi.Context(local_rank, stream=stream)
    t0 = time.perf_counter()
    if device_setup:
        ctx.set_problem_on_device(P)
    else:
        ctx.set_problem(P)
    t_upload = time.perf_counter() - t0
    comm_used = "none"
    if world > 1:
        comm_used = args.comm
        if args.comm == "peer":
            # CUDA IPC can be unavailable (container without shared IPC namespace, no P2P): every
            # rank must take the same path, so agree on the outcome before falling back to NCCL.
            try:
                pt.dist.connect_peers(ctx, P, dist, rank, world)
                ok_local = 1.0
            except Exception as e:  # noqa: BLE001
                ok_local = 0.0
                print(f"rank {rank}: peer-memory setup failed: {e}", file=sys.stderr)
            t = torch.tensor([ok_local], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            if float(t.item()) < 1.0:
                if rank == 0:
                    print("peer-memory setup failed on some rank, using NCCL", file=sys.stderr)
                comm_used = "nccl"
        if comm_used == "nccl":
            pt.dist.init_nccl(ctx, abi, dist, rank, world)   # ptb_comm_init drops any peer state
    ndofs_global = P.n_global * P.bs
    if world > 1 and comm_used == "peer":
        # the persistent CG loop pays below ~2 M DOFs per GPU; every rank must
        # take the same path, so the decision comes from the global size
        ctx.set_cg_persistent(1 if P.bs == 3 and ndofs_global / world <= 3_000_000 else 0)
    nnz_local = ctx.nnz if device_setup else P.nnz
    nnz_global = allsum(float(nnz_local * P.bs * P.bs))

    # pinned host buffers for the e2e leg (host-resident inputs of the hot path)
    nl = (P.n_owned + P.n_ghost) * P.bs
    if device_setup:
        x_host, _ = ctx.mesh(topology=False)
        f_host, g_host = ctx.source()
    else:
        x_host, f_host = np.array(P["x"]), np.array(P["f"])
        g_host = np.array(P["g"]) if len(P["g"]) else None
    x_pin = torch.from_numpy(x_host).pin_memory()
    f_pin = torch.from_numpy(f_host).pin_memory()
    g_pin = torch.from_numpy(g_host).pin_memory() if g_host is not None else None
    b_pin = torch.empty(P.n_owned * P.bs, dtype=torch.float64).pin_memory()
    u_pin = torch.empty(nl, dtype=torch.float64).pin_memory()
    h2d = x_pin.numel() * 8 + f_pin.numel() * 8 + (g_pin.numel() * 8 if g_pin is not None else 0)
    d2h = b_pin.numel() * 8 + u_pin.numel() * 8

    state = {}

    def step(e2e=False):
        if e2e:
            ctx.update_geometry(x_pin.numpy())
            ctx.set_source(f_pin.numpy(), None if g_pin is None else g_pin.numpy())
        ctx.assemble_matrix()
        ctx.assemble_vector()
        k, rel = ctx.cg_solve(kmax=args.kmax, rtol=1e-8, precond="jacobi")
        if e2e:
            ctx.rhs(out=b_pin.numpy())
            ctx.solution(out=u_pin.numpy())
        state.update(k=k, rel=rel, am=ctx.stage_ms(abi.STAGE_ASSEMBLE_MATRIX),
                     av=ctx.stage_ms(abi.STAGE_ASSEMBLE_VECTOR), sv=ctx.stage_ms(abi.STAGE_SOLVE))
        return k

    for _ in range(warmup):
        step()

    def timed(nsteps, e2e):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        am = sv = 0.0
        iters = 0
        barrier()
        ev0.record()
        for _ in range(nsteps):
            iters += step(e2e)
            am += state["am"]
            sv += state["sv"]
        ev1.record()
        barrier()
        ms = allmax(ev0.elapsed_time(ev1))
        return ms, allmax(am), allmax(sv), iters

    launches0 = ctx.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, am_ms, sv_ms, iters = timed(steps, e2e=False)
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    # the end-to-end leg repeats whole steps with host buffers; bounded so that a long --steps run
    # of the 2.5 s elasticity step still ends within minutes (the rate does not depend on the count)
    steps_e2e = min(steps, 5)
    ms_e2e, am_e2e, sv_e2e, iters_e2e = timed(steps_e2e, e2e=True)

    value = iters * ndofs_global / (sv_ms * 1e-3)
    nnz_per_s = nnz_global * steps / (am_ms * 1e-3)
    # e2e: (iterations * DOFs) / whole time of the host-buffer steps -- copies and assembly included
    e2e_value = iters_e2e * ndofs_global / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel (SpMV inside CG), timed live with CUDA events on this
    # rank's stream: the plain operator kernel on resident data (no halo wait, no reduction
    # publish), so the number is this GPU's kernel time at any N; max over ranks is reported ----
    step()  # leaves p, r, x in their end-of-solve state
    t_spmv = allmax(ctx.time_kernel(abi.KERNEL_SPMV, 30))
    t_upd = allmax(ctx.time_kernel(abi.KERNEL_CG_UPDATE, 30))
    t_dir = allmax(ctx.time_kernel(abi.KERNEL_CG_DIRECTION, 30))
    t_am = allmax(ctx.time_kernel(abi.KERNEL_ASSEMBLE_MATRIX, 5))
    t_av = allmax(ctx.time_kernel(abi.KERNEL_ASSEMBLE_VECTOR, 5))
    n, bs = P.n_owned, P.bs
    spmv_b = 12 * nnz_local + 20 * n if bs == 1 else 76 * nnz_local + 52 * n
    cg_b = spmv_b + 96 * n * bs
    asm_b = 8 * nnz_local * bs * bs + P.n_cells * (16 + 4 * P.nd) + 25 * (n + P.n_ghost)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = spmv_b / (t_spmv * 1e-3) / 1e9
    # ncu traffic belongs to one capture: only quoted when this run has the capture's configuration
    traffic, traffic_src = None, None
    switches = {k: v for k, v in os.environ.items() if k.startswith("PTB_")}
    tpath = os.path.join(ROOT, "profiles", "spmv_traffic.json")
    if os.path.exists(tpath) and world == 1 and args.ndofs is None and not switches:
        try:
            tj = json.load(open(tpath)).get(wl_name, {})
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
        except Exception:
            pass
    it_ms = sv_ms / max(iters, 1)
    roofline = {"bound": "hbm", "kernel": f"spmv_sell<{bs}> (y = A p + p.y inside CG)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": spmv_b, "ms_per_launch": t_spmv,
                "cols_explicit_fraction": ctx.cols_explicit_fraction(),
                "spmv_stored_entries": ctx.spmv_stored_entries(),
                "pattern_entries": nnz_local,
                "note": ("per-rank kernel time (max over ranks), launched back to back on resident data; "
                         "achieved counts the ALGORITHMIC CSR bytes (8 B value + 4 B column per nnz, "
                         "76 B per 3x3 block); scalar matrices store one column delta per 32 rows where "
                         "the stencil is translation invariant (cols_explicit_fraction), so frac can "
                         "exceed 1 there -- block matrices store every column index"),
                "other_kernels": {
                    "cg_update": {"ms": t_upd, "GBps": 32.0 * n * bs / t_upd / 1e6},
                    "cg_direction": {"ms": t_dir, "GBps": 48.0 * n * bs / t_dir / 1e6},
                    "assemble_matrix": {"ms": t_am, "GBps": asm_b / t_am / 1e6,
                                        "frac_of_hbm_peak": asm_b / t_am / 1e6 / peak,
                                        "nnz_per_s": nnz_local * bs * bs / (t_am * 1e-3)},
                    "assemble_vector": {"ms": t_av}},
                "cg_iteration": {"ms_measured": it_ms, "ms_sum_of_kernels": t_spmv + t_upd + t_dir,
                                 "overhead_frac": it_ms / (t_spmv + t_upd + t_dir) - 1.0,
                                 "frac_of_hbm_roofline": cg_b / (it_ms * 1e-3) / 1e9 / peak}}

    cpu = cpu_baseline(pt, args, wl_name) if (with_cpu and rank == 0 and world == 1) else None

    out = None
    if rank == 0:
        wl = WORKLOADS[wl_name]
        walk3 = os.environ.get("PTB_ASM_WALK3", "1") != "0"
        ring = os.environ.get("PTB_ASM_RING", "1") != "0"   # edge rings: host- or device-built maps
        walk = os.environ.get("PTB_ASM_WALK", "1") != "0"
        out = {
            "metric": "cg_dof_iters_per_s", "value": value, "unit": "DOF-iters/s",
            "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": wl[1],
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "assembled_nnz_per_s": nnz_per_s,
            "stage_ms": {"assemble_matrix": am_ms / steps, "assemble_vector": state["av"],
                         "solve": sv_ms / steps},
            "cg_iterations": iters // steps, "rel_residual": state["rel"],
            "ndofs_global": ndofs_global, "nnz_global": nnz_global,
            "config": {"workload": wl[4], "problem_type": ptype, "order": order,
                       "ndofs_arg": ndofs_arg, "scaling_type": scaling,
                       "base_box": list(base[:3]), "refinements": base[3],
                       "fine_box": list(dims), "partition": f"z-slabs x{world}",
                       "comm": ("none" if world == 1 else
                                "nvlink peer memory (halo pull + window all-reduce in the CG kernels)"
                                if comm_used == "peer" else "nccl send/recv + allreduce"),
                       "l2": "inputs larger than L2 (matrix + vectors >> 126 MB per GPU)"
                             if nnz_local * 12 * bs > 3e8 else
                             "per-GPU working set near L2 size (strong scaling / small config)",
                       "refined_mesh_note": "r>0 generated as the (N<<r) box directly, not by "
                                            "Plaza refinement (same entity counts)" if base[3] else None,
                       "switches": switches,
                       "matrix_kernel": ("assemble_matrix_pk_binned" if order > 1 else
                                         ("assemble_matrix_p1_ring3 (edge rings)" if ring else
                                          "assemble_matrix_p1_walk3 (star walk)" if walk3 else
                                          "assemble_matrix_p1<3> (cell order)") if ptype == "elasticity" else
                                         "assemble_matrix_p1_walk (star walk)" if walk else
                                         "assemble_matrix_p1<1> (cell order)")},
            "e2e": {"value": e2e_value, "unit": "DOF-iters/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / steps_e2e, "steps": steps_e2e,
                    "note": "per step: x, f, g host->device from pinned memory, assemble A and b, "
                            "solve, b and u device->host; value = iterations*DOFs / whole e2e time"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "setup_s": {"host_mesh_dofmap_pattern": t_host, "slot_map_and_upload": t_upload,
                        "device_setup": device_setup},
            "device_bytes": ctx.device_bytes(),
        }
    barrier()
    ctx.close()
    del ctx, P
    if out is not None and world == 1 and not device_setup and not args.no_renumbered:
        # the shuffle of a 20 M-row secondary problem costs 27 s of host time for one kernel timing:
        # the secondary block carries the realistic (rcm) numbering only
        kinds = ("rcm",) if secondary else ("rcm", "random")
        out["renumbered"] = renumbered_numbers(pt, wl_name, args, peak, kinds)
    return out


def renumbered_numbers(pt, wl_name, args, peak, kinds=("rcm", "random")):
    """The same workload with the owned dofs renumbered the way a DOLFINx dofmap is (not lattice-
    lexicographic): "rcm" = reverse Cuthill-McKee (banded, local, no translation invariance -- the
    realistic case: every column index of the scalar operator goes explicit), "random" = seeded
    shuffle (no locality: the worst case for the gather of p). One solve for rcm, kernel timings
    for both; N = 1 only (the stand-in renumbers single-rank problems)."""
    import torch
    abi = pt.abi
    res = {}
    ptype, order, dims, base, scaling, ndofs_arg = sizing(pt, wl_name, 1, args.ndofs)
    for kind in kinds:
        t0 = time.perf_counter()
        P = pt.host.Problem(ptype, order, *dims, renumber=kind, seed=1)
        ctx = abi.Context(torch.cuda.current_device(), stream=torch.cuda.current_stream().cuda_stream)
        ctx.set_problem(P)
        t_setup = time.perf_counter() - t0
        ctx.assemble_matrix()
        ctx.assemble_vector()
        n, bs, nnz = P.n_owned, P.bs, P.nnz
        r = {"setup_s": t_setup, "cols_explicit_fraction": ctx.cols_explicit_fraction(),
             "stage_ms": {"assemble_matrix": ctx.stage_ms(abi.STAGE_ASSEMBLE_MATRIX),
                          "assemble_vector": ctx.stage_ms(abi.STAGE_ASSEMBLE_VECTOR)}}
        if kind == "rcm":
            k, rel = ctx.cg_solve(kmax=args.kmax, rtol=1e-8, precond="jacobi")
            sv = ctx.stage_ms(abi.STAGE_SOLVE)
            r.update(cg_iterations=k, rel_residual=rel, value=k * n * bs / (sv * 1e-3),
                     unit="DOF-iters/s", solve_ms=sv)
        else:
            ctx.cg_solve(kmax=5, rtol=1e-8, precond="jacobi")   # leaves p, r, x populated
        t_spmv = ctx.time_kernel(abi.KERNEL_SPMV, 20)
        spmv_b = 12 * nnz + 20 * n if bs == 1 else 76 * nnz + 52 * n
        r.update(spmv_ms=t_spmv, spmv_GBps=spmv_b / t_spmv / 1e6, spmv_frac=spmv_b / t_spmv / 1e6 / peak)
        res[kind] = r
        ctx.close()
        del ctx, P
    return res


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    pt = importlib.import_module("performance-test_b200")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    env = dict(world=world, rank=rank, local_rank=local_rank, dist=dist, barrier=barrier,
               allmax=allmax, allsum=allsum)
    head = args.workload or DEFAULT_HEADLINE
    # the CPU arm runs on the GPU arm's own mesh: at 100 M DOFs (configs[4]) that is minutes of host time
    # for one line, so that workload carries no cpu_baseline
    line = measure(pt, env, head, args, args.steps, args.warmup,
                   with_cpu=not args.no_cpu_baseline and head != "elasticity_weak")
    if args.workload is None and not args.no_secondary:
        # the weak-scaling config of BASELINE.json in the same run: same keys, under "secondary"
        sec = measure(pt, env, DEFAULT_SECONDARY, args, min(args.steps, 3), min(args.warmup, 3),
                      with_cpu=False, secondary=True)
        if rank == 0:
            line["secondary"] = sec
    if rank == 0:
        print(json.dumps(line), flush=True)
    barrier()
    if world > 1:
        dist.destroy_process_group()


def _json_only_stdout():
    """The contract is ONE JSON line on stdout. Libraries loaded below print banners of their own
    (e.g. "NCCL version ..." from ncclCommInitRank writes straight to file descriptor 1), so the
    process's stdout is pointed at stderr for the duration of the run and `print` keeps the real
    stdout for the JSON line alone."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)


if __name__ == "__main__":
    _json_only_stdout()
    main()
