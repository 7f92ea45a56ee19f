#!/usr/bin/env python
"""bench.py — GMRES-IR time-to-solution / inner iterations per second on the B200 backend.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cd27:256]

One "step" = one complete GMRES-IR solve (x0 = 0 -> the reference's stopping rule) of the BASELINE.json workload:
synthetic 3-D 27-point convection-diffusion, 256^3 = 16.7 M rows, 449 M nonzeros, GMRES-IR(100), CGS2, identity
preconditioner, tol 1e-6 (configs[2], the configuration the metric is quoted on; it fits one GPU).
Prints ONE JSON line (rank 0).  `value` = inner iterations of all timed solves / device time (CUDA events, inputs
resident in HBM); `e2e` = the same through the host-buffer C-ABI call (H2D + plan + solve + D2H per step);
`roofline` = algorithmic bytes / CUDA-event time of the dominant kernel class, measured inside the timed region.
For N > 1 the driver launches this file under torchrun; ranks own 1-D row blocks.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gmres_ir_inner_iterations_per_sec"
UNIT = "it/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cd27:256")
    ap.add_argument("--rlen", type=int, default=100)
    ap.add_argument("--tol", type=float, default=1e-6)
    ap.add_argument("--orth", default="cgsr")
    ap.add_argument("--mode", default="mixed", choices=["mixed", "baseline", "single-prec", "single"])
    ap.add_argument("--max-restarts", type=int, default=1000)
    ap.add_argument("--cpu-sample", default="auto", help="CPU baseline workload of the b200 arm (auto: the bench workload itself, one complete solve; "
                    "anything else is labelled as not being the bench workload and yields no parity block)")
    ap.add_argument("--sweep-matrix", default="cd27:128", help="--workload sweep: matrix of the kernel sweep (both arms); beyond the host LLC by default")
    ap.add_argument("--partition", default="rows", choices=["rows", "nnz"], help="N > 1: equal row blocks or nnz-balanced split points (SURVEY.md §8e)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-multi-restart", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--tune", action="append", default=[], metavar="KEY=INT", help="mpg_set_tuning knob (experiments), e.g. --tune use_pdl=0")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(",") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, pw, reasons = [], [], [], set()
        for r in rows:
            try:
                r = [c.strip() for c in r]
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm), "reasons": sorted(reasons)}


_HOST_THREADS = None


def host_threads():
    """host cores this process may use (affinity-aware).  Read ONCE, before the first OpenMP region: with OMP_PROC_BIND the
    runtime later pins the calling thread to a single core and sched_getaffinity would then say 1."""
    global _HOST_THREADS
    if _HOST_THREADS is None and os.environ.get("MPG_HOST_THREADS"):
        _HOST_THREADS = int(os.environ["MPG_HOST_THREADS"])   # measured before any OpenMP runtime pinned this thread (also survives re-imports)
    if _HOST_THREADS is None:
        try:
            _HOST_THREADS = max(1, len(os.sched_getaffinity(0)))
        except Exception:
            _HOST_THREADS = max(1, os.cpu_count() or 1)
    return _HOST_THREADS


def force_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must use every host core it can.  Call BEFORE the first
    import of numpy / torch in this process (libgomp reads the environment once) - torch.set_num_threads covers the rest."""
    nt = host_threads()
    os.environ["MPG_HOST_THREADS"] = str(nt)
    for k in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = str(nt)
    os.environ.setdefault("OMP_PROC_BIND", "spread")    # automated.py:13-15
    os.environ.setdefault("OMP_PLACES", "threads")
    os.environ["MKL_DYNAMIC"] = "FALSE"
    return nt


def cpu_reference_run(workload, rlen, tol, orth, max_restarts, mode="mixed", want_hist=True):
    """ONE solve of `workload` (the real thing, never a scaled stand-in) by the reference's own CPU implementation of the path on
    this box's host cores: oracle/_ref (the reference's sources built against the MKL inside libtorch) when present, else the
    oracle port.  Returns counts, seconds inside the reference's timing window (gmres_perf_test.cpp:165-167), the per-iteration
    history and the post-solve fp64 resNorm / errNorm (gmres_perf_test.cpp:169-175)."""
    nt = force_host_threads()
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle as orc
    try:
        import oracle_ref
        have_ref = oracle_ref.available()
    except Exception:
        have_ref = False
    import torch
    torch.set_num_threads(nt)   # omp_set_num_threads + mkl_set_num_threads for this thread, whatever the environment said at start-up
    t_setup = time.perf_counter()
    rm, ind, val = orc.gen(workload)
    n = len(rm) - 1
    xt = orc.rand_vect(n, 42)
    b = np.zeros(n)
    orc.spmv(rm, ind, val, 1.0, xt, 0.0, b)
    t_setup = time.perf_counter() - t_setup
    t0 = time.perf_counter()
    if have_ref:
        r = oracle_ref.gmres(rm, ind, val, b, true_x=xt, mode=mode, orth=orth, rlen=rlen, tol=tol, max_restarts=max_restarts)
        kind, cores = "reference", int(oracle_ref.num_threads())
        res_norm, err_norm = float(r["res_norm"]), float(r["err_norm"])
    else:
        r = orc.gmres(rm, ind, val, b, mode=mode, orth=orth, rlen=rlen, tol=tol, max_restarts=max_restarts)
        kind, cores = "port", int(orc.num_threads())
        res = b.copy(); orc.spmv(rm, ind, val, -1.0, r["x"], 1.0, res)
        res_norm, err_norm = float(np.linalg.norm(res)), float(np.linalg.norm(r["x"] - xt))
    dt = float(r.get("gmres_seconds", time.perf_counter() - t0))
    out = {"workload": workload, "n_rows": int(n), "nnz": int(len(ind)), "kind": kind, "cores": cores, "seconds": dt, "setup_seconds": t_setup,
           "iters": int(r["total_iters"]), "restarts": int(r["total_restarts"]), "status": int(r["status"]),
           "resNorm": res_norm, "errNorm": err_norm, "b_norm": float(np.linalg.norm(b))}
    if want_hist:
        out["hist_inner"] = np.asarray(r["hist_inner"], dtype=np.float64)
        out["hist_outer"] = np.asarray(r["hist_outer"], dtype=np.float64)
    return out


def parity_block(cpu, gpu_hist, gpu_iters, gpu_restarts, gpu_status, gpu_res, gpu_err):
    """reference-vs-GPU comparison on the SAME inputs at the bench size (SURVEY.md §8d "parity acceptance"): counts, the
    per-iteration |s(k+1)|/||M^-1 b|| history, post-solve fp64 norms."""
    import numpy as np
    hr = np.asarray(cpu["hist_inner"]); hg = np.asarray(gpu_hist)
    m = int(min(len(hr), len(hg)))
    dev = None
    if m:
        live = hr[:m] > 1e-4 * hr[0]    # entries below 1e-4 of the first are rounding noise in fp32
        rel = np.abs(hg[:m] - hr[:m]) / np.maximum(np.abs(hr[:m]), 1e-300)
        dev = float(rel[live].max()) if live.any() else 0.0
    rd = lambda a, b: abs(a - b) / max(abs(b), 1e-300)
    return {"against": f"{cpu['kind']} ({'oracle/_ref: the reference sources + MKL' if cpu['kind'] == 'reference' else 'oracle port'}), {cpu['cores']} host threads, "
                       f"{cpu['workload']} full size, one solve in {cpu['seconds']:.2f} s",
            "iters": [int(gpu_iters), cpu["iters"]], "restarts": [int(gpu_restarts), cpu["restarts"]], "status": [int(gpu_status), cpu["status"]],
            "iters_equal": int(gpu_iters) == cpu["iters"], "restarts_equal": int(gpu_restarts) == cpu["restarts"],
            "hist_len_compared": m, "hist_dev": dev,
            "resNorm": [gpu_res, cpu["resNorm"]], "resNorm_rel_diff": rd(gpu_res, cpu["resNorm"]),
            "errNorm": [gpu_err, cpu["errNorm"]], "errNorm_rel_diff": rd(gpu_err, cpu["errNorm"]),
            "order": "[this backend, reference]"}


CPU_MAX_STEPS = 2   # a cd27:256 solve is ~15-30 s on the host cores: the CPU arm times at most this many real solves, no warm-up repeats


def run_reference(args, rank, world):
    if rank != 0:
        return
    if args.workload == "sweep":
        import sweep
        sweep.run_reference(args)
        return
    steps = max(1, min(args.steps, CPU_MAX_STEPS))
    total_it, total_t = 0, 0.0
    last = None
    for i in range(steps):
        last = cpu_reference_run(args.workload, args.rlen, args.tol, args.orth, args.max_restarts, args.mode, want_hist=False)
        total_it += last["iters"]; total_t += last["seconds"]
    assert last["cores"] > 1 or host_threads() == 1, f"CPU arm is running on {last['cores']} thread(s) of {host_threads()}"
    value = total_it / total_t
    sample = (f"{args.workload} at full size ({last['n_rows']} rows, {last['nnz']} nonzeros), {steps} complete GMRES-IR({args.rlen}) solve(s), "
              f"{last['iters']} inner iterations / {last['restarts']} restart checks each, timed inside the reference's own window "
              f"(gmres_perf_test.cpp:165-167); no warm-up repeats, at most {CPU_MAX_STEPS} steps whatever --steps asks for")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": 0,
            "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": 1e3 * total_t / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 inner / f64 outer (CPU)" if args.mode == "mixed" else args.mode, "data": "synthetic",
            "config": {"workload": args.workload, "mode": args.mode, "n_rows": last["n_rows"], "nnz": last["nnz"], "restart_length": args.rlen, "tol": args.tol,
                       "orth": args.orth, "prec": "identity", "iters_per_solve": last["iters"], "restarts_per_solve": last["restarts"],
                       "status": last["status"], "time_to_solution_s": total_t / steps, "resNorm": last["resNorm"], "errNorm": last["errNorm"],
                       "rel_res": last["resNorm"] / last["b_norm"], "setup_seconds_untimed": round(last["setup_seconds"], 2)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": last["cores"], "kind": last["kind"], "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        force_host_threads()
        run_reference(args, rank, world)
        return
    if args.workload == "sweep":
        import sweep
        sweep.run_b200(args, rank, world, local_rank)
        return

    import numpy as np
    import torch
    import gmres_b200 as g

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import bench_dist
        bench_dist.main(args, rank, world, local_rank)
        return

    dev = f"cuda:{local_rank}"
    ctx = g.Context(local_rank)
    for kv in args.tune:
        ctx.set_tuning(kv.split("=")[0], int(kv.split("=")[1]))
    peak, peak_src = measured_peaks()

    # ---- problem (device-resident; construction is outside the timed region like gmres_perf_test.cpp:408-421) ----
    rm, ind, val = ctx.gen(args.workload)
    n, nnz = rm.numel() - 1, ind.numel()
    A = g.CSR(ctx, rm, ind)
    xt = torch.from_numpy(ctx.rand_vect(n, 42)).to(dev)
    b = torch.zeros(n, dtype=torch.float64, device=dev)
    ctx.spmv(A, val, 1.0, xt, 0.0, b)
    val32 = torch.empty(nnz, dtype=torch.float32, device=dev)
    ctx.copy(val, val32)  # SparseMatrix<float>(A): the reference's "prec" window, gmres_perf_test.cpp:135-136
    x = torch.zeros(n, dtype=torch.float64, device=dev)
    kw = dict(mode=args.mode, orth=args.orth, conv="base", prec="identity", rlen=args.rlen, tol=args.tol, max_restarts=args.max_restarts)

    def solve():
        x.zero_()
        return ctx.gmres(A, val, b, x, vals32=val32, hist_cap=1, **kw)

    for _ in range(args.warmup):
        r = solve()
    torch.cuda.synchronize()

    # working set of one solve: matrix (fp32 + fp64 values, indices) + Krylov basis.  When it does not dwarf the 126 MB
    # L2, L2 is flushed (a 512 MB buffer is overwritten) between timed steps and each step is timed with its own events.
    ws_bytes = nnz * 16 + (n + 1) * 4 + n * (args.rlen + 1) * (8 if args.mode in ("baseline", "single-prec") else 4)
    flush = ws_bytes < 8 * 126e6
    flush_buf = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev) if flush else None

    # ---- timed region: K solves, device time from CUDA events.  The per-kernel-class timers (CUDA events around every launch, on
    # the launching stream) run in a SECOND pass over the same K solves right after it: kernel durations are the same, the step
    # is not perturbed by ~1000 event records, and N = 1 is measured exactly like N = 2/4/8 ----
    prof_in_region = False   # same rule at every N (bench_dist.py): event records between launches cost ~1.7 % of the step, so the timed region carries none
    ctx.prof_enable(prof_in_region)
    ctx.prof_reset()
    sampler = ClockSampler(local_rank)
    launches0 = ctx.launches()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    iters = restarts = 0
    status = 1
    for i in range(args.steps):
        if flush:
            flush_buf.fill_(i & 0xFF)
        evs[i][0].record()
        r = solve()
        evs[i][1].record()
        iters += r["total_iters"]; restarts += r["total_restarts"]; status = r["status"]
    torch.cuda.synchronize()
    clocks = sampler.stop()
    step_ms = sorted(a.elapsed_time(b) for a, b in evs)
    total_ms = sum(step_ms)
    launches = ctx.launches() - launches0
    profiled_ms = None
    if not prof_in_region:
        ctx.prof_enable(True); ctx.prof_reset()
        pe = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        torch.cuda.synchronize()
        pe[0].record()
        for i in range(args.steps):
            if flush:
                flush_buf.fill_(i & 0xFF)
            solve()
        pe[1].record()
        torch.cuda.synchronize()
        profiled_ms = pe[0].elapsed_time(pe[1]) / args.steps
    prof = ctx.prof_get()
    ctx.prof_enable(False)

    # the reference's own timing window (gmres_perf_test.cpp:165-167) also covers allocating and zero-filling the
    # workspace inside gmres_singleUpdate (gmres.cpp:147-157): one solve on a fresh context = workspace + plan + packed copy
    # allocated inside the window
    incl_alloc_ms = None
    try:
        ctx2 = g.Context(local_rank)
        A2 = g.CSR(ctx2, rm, ind)
        x2 = torch.zeros(n, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx2.gmres(A2, val, b, x2, vals32=val32, hist_cap=1, **kw)
        torch.cuda.synchronize()
        incl_alloc_ms = (time.perf_counter() - t0) * 1e3
        del A2, x2
        ctx2.close()
    except Exception as e:   # reported, never fatal for the bench line
        incl_alloc_ms = f"failed: {e}"

    # a multi-restart solve next to the headline one (tol 1e-6 stops after ONE cycle on this system: update, second residual and
    # re-entry are 1.5 % of what is timed above): same system, tol 1e-12, device time of one solve after one warm solve
    multi = None
    if not args.no_multi_restart:
        kw12 = dict(kw, tol=1e-12, max_restarts=50)
        x.zero_(); ctx.gmres(A, val, b, x, vals32=val32, hist_cap=1, **kw12)
        x.zero_()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record(); rm12 = ctx.gmres(A, val, b, x, vals32=val32, hist_cap=1, **kw12); m1.record()
        torch.cuda.synchronize()
        res12 = b.clone(); ctx.spmv(A, val, -1.0, x, 1.0, res12)
        multi = {"tol": 1e-12, "status": int(rm12["status"]), "iters": int(rm12["total_iters"]), "restart_checks": int(rm12["total_restarts"]),
                 "ms": round(m0.elapsed_time(m1), 3), "it_per_s": round(rm12["total_iters"] / (m0.elapsed_time(m1) * 1e-3), 1),
                 "rel_res": ctx.nrm2(res12) / ctx.nrm2(b)}
        x.zero_(); r = solve()   # leave x = the headline solution for the checks below

    # post-solve check the reference prints (gmres_perf_test.cpp:169-178)
    res = b.clone(); ctx.spmv(A, val, -1.0, x, 1.0, res)
    res_norm = ctx.nrm2(res)
    err = x - xt
    err_norm = ctx.nrm2(err)
    b_norm = ctx.nrm2(b)

    value = iters / (total_ms * 1e-3)
    kernels = {}
    for name, p in prof.items():
        if p["launches"] == 0:
            continue
        gbs = p["bytes"] / (p["ms"] * 1e-3) / 1e9 if p["ms"] > 0 else 0.0
        kernels[name] = {"ms_total": round(p["ms"], 3), "share": round(p["ms"] / total_ms, 4), "launches": p["launches"],
                         "avg_ms": round(p["ms"] / p["launches"], 4), "algorithmic_GB_per_launch": round(p["bytes"] / p["launches"] / 1e9, 5),
                         "achieved_GBps": round(gbs, 1), "frac_of_peak": round(gbs / peak, 4)}
    dom = max((k for k in kernels if k in ("vpass", "spmv_f32", "gemvn", "spmv_f64", "gemvt")), key=lambda k: kernels[k]["ms_total"])
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_GBps"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["frac_of_peak"], "traffic": None, "peak_source": peak_src,
                "note": "achieved = algorithmic bytes (SURVEY.md §8d / DESIGN.md §4) of all launches of this kernel class in the timed region / "
                        "their CUDA-event time on the launching stream"}
    prof_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof_path) and args.workload == "cd27:256":
        try:
            roofline["traffic"] = json.load(open(prof_path)).get(dom)
        except Exception:
            pass

    # ---- end to end through the host-buffer C-ABI entry point (pinned host memory) ----
    e2e = None
    if not args.no_e2e:
        h_rm = torch.empty(rm.shape, dtype=rm.dtype, pin_memory=True); h_rm.copy_(rm)
        h_ind = torch.empty(ind.shape, dtype=ind.dtype, pin_memory=True); h_ind.copy_(ind)
        h_val = torch.empty(val.shape, dtype=val.dtype, pin_memory=True); h_val.copy_(val)
        h_b = torch.empty(b.shape, dtype=b.dtype, pin_memory=True); h_b.copy_(b)
        h_x = torch.zeros(n, dtype=torch.float64).pin_memory()
        torch.cuda.synchronize()
        h_x.zero_(); ctx.gmres_host(h_rm, h_ind, h_val, h_b, h_x, hist_cap=1, **kw)  # warm
        t0 = time.perf_counter()
        it2 = 0
        for _ in range(args.e2e_steps):
            h_x.zero_()
            r2 = ctx.gmres_host(h_rm, h_ind, h_val, h_b, h_x, hist_cap=1, **kw)
            it2 += r2["total_iters"]
        dt = time.perf_counter() - t0
        h2d_inputs = h_rm.numel() * 4 + h_ind.numel() * 4 + h_val.numel() * 8 + 2 * n * 8
        # bytes that actually crossed the link, counted by the library from the copies it issued: with host_overlap the fp32 values the
        # host cores cast travel too (4 B per nonzero on top of the caller's inputs)
        e2e = {"value": it2 / dt, "unit": UNIT, "h2d_bytes_per_step": int(r2["h2d_bytes"]), "d2h_bytes_per_step": int(n * 8),
               "input_bytes_per_step": int(h2d_inputs), "ms_per_step": 1e3 * dt / args.e2e_steps, "steps": args.e2e_steps,
               "h2d_ms_until_solve_starts": r2["h2d_ms"], "h2d_ms_until_last_byte": r2["h2d_all_ms"], "solve_ms": r2["solve_ms"], "d2h_ms": r2["d2h_ms"],
               "host_cast_threads": int(r2["host_overlap"]),
               "call": "mpg_gmres_solve_host (pinned host CSR + b in, x out)" + (
                   "; overlapped: indices first while host threads cast the values to fp32, solve starts on the fp32 operator, fp64 values land during "
                   "the first restart cycle" if r2["host_overlap"] else "; serial H2D")}
        del h_rm, h_ind, h_val, h_b

    # ---- CPU baseline on this box's host cores: ONE complete solve of the same workload by the reference's own CPU path, and the
    # parity block it yields at the bench size (same inputs: both sides generate the workload from the same frozen definition) ----
    cpu = parity = None
    if not args.no_cpu_baseline:
        sample_wl = args.workload if args.cpu_sample == "auto" else args.cpu_sample
        c = cpu_reference_run(sample_wl, args.rlen, args.tol, args.orth, args.max_restarts, args.mode)
        cpu = {"value": c["iters"] / c["seconds"], "unit": UNIT, "cores": c["cores"], "kind": c["kind"],
               "sample": (f"{sample_wl}{'' if sample_wl == args.workload else ' (NOT the bench workload ' + args.workload + '; unscaled)'}: one complete "
                          f"GMRES-IR({args.rlen}) solve at full size, {c['iters']} inner iterations / {c['restarts']} restart checks in {c['seconds']:.2f} s "
                          f"inside the reference's timing window (gmres_perf_test.cpp:165-167)"),
               "seconds": c["seconds"], "iters": c["iters"], "restarts": c["restarts"], "resNorm": c["resNorm"], "errNorm": c["errNorm"]}
        if sample_wl == args.workload:
            x.zero_()
            rp = ctx.gmres(A, val, b, x, vals32=val32, **kw)   # one more (untimed) solve with the full residual history
            res = b.clone(); ctx.spmv(A, val, -1.0, x, 1.0, res)
            parity = parity_block(c, rp["hist_inner"], rp["total_iters"], rp["total_restarts"], rp["status"], ctx.nrm2(res), ctx.nrm2(x - xt))

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"mixed": "f32 inner / f64 outer", "baseline": "f64", "single-prec": "f64 (f32 preconditioner)", "single": "f32"}[args.mode],
            "data": "synthetic",
            "config": {"workload": args.workload, "mode": args.mode, "n_rows": n, "nnz": nnz, "restart_length": args.rlen, "tol": args.tol, "orth": args.orth,
                       "prec": "identity", "iters_per_solve": iters // max(args.steps, 1), "restarts_per_solve": restarts // max(args.steps, 1),
                       "time_to_solution_s": total_ms * 1e-3 / args.steps, "status": int(status),
                       "step_ms_min_med_max": [round(step_ms[0], 3), round(step_ms[len(step_ms) // 2], 3), round(step_ms[-1], 3)],
                       "first_solve_incl_workspace_alloc_ms": incl_alloc_ms if not isinstance(incl_alloc_ms, float) else round(incl_alloc_ms, 2),
                       "resNorm": res_norm, "errNorm": err_norm, "rel_res": res_norm / b_norm, "multi_restart": multi,
                       "l2": (f"working set {ws_bytes / 1e9:.2f} GB: L2 flushed (512 MB overwrite) between timed steps" if flush else
                              f"working set {ws_bytes / 1e9:.1f} GB (matrix + basis) >> 126 MB L2; no flush needed"),
                       "kernel_timers": ("CUDA events around every launch inside the timed region" if prof_in_region else
                                         "CUDA events around every launch in a second pass over the same K solves right after the timed region (same rule at "
                                         "every N; the event records would perturb the timed step, kernel durations are unaffected)"),
                       "profiled_pass_ms_per_step": None if profiled_ms is None else round(profiled_ms, 3)},
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu, "parity": parity, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
