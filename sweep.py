"""sweep.py — BASELINE.json configs[4] behind `bench.py --workload sweep`: the kernel_perf_test sweep (kernel_perf_test.cpp:84-137,
170-179 is the shape: ops x {float, double}, one warm pass then timed passes, basis / vectors from mt19937 floats, --vcols).

  * SpMV fp32 vs fp64 (y = A x, beta = 0);
  * one Arnoldi orthogonalisation step GS::add_vector (orthogonalise + norm + normalise, Orthogonalization.hpp:51-60) with
    MGS / CGS / CGS2 at basis width k + 1 = m and averaged over a whole restart cycle k = 0 .. m - 1, for m in {25, 50, 100}.

b200 arm: CUDA events around every call, median of `trials`, on --sweep-matrix (default cd27:128, the CPU arm's size - beyond the
host LLC) and on cd27:256 (the headline size).  reference arm (`--impl reference`): the SAME ops through the reference's own
kernels_mkl.cpp + Orthogonalization.hpp built in oracle/_ref (spmv<T,MKL>, GS<T,Kernel<T,MKL>,MKL>::add_vector), all host threads,
warm pass + 3 timed passes.  Both print ONE JSON line; `value` = CGS2 steps of the m = 100 cycle + SpMV per second in fp32 (the
kernel pair of one GMRES-IR inner iteration), so the driver's ratio is like for like."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
METRIC = "arnoldi_kernel_pair_fp32_cgs2_m100_per_sec"
UNIT = "steps/s"
MS = (25, 50, 100)
ORTHS = ("mgs", "cgs", "cgsr")


def ks_for(m, dense):
    """values of k timed for the cycle average of restart length m (every k on the GPU; a uniform subsample on the CPU)"""
    step = 1 if dense else (1 if m <= 25 else (2 if m <= 50 else 4))
    return list(range(0, m, step))


def bytes_add_vector(orth, k1, n, s):
    """algorithmic bytes of one add_vector: orthogonalisation + normalise 2 n s"""
    if orth == "mgs":
        return 4.0 * k1 * n * s + 2.0 * n * s                # pairwise-fused MGS (the unfused loop moves 5 k1 n s): fusion must not inflate GB/s
    if orth == "cgs":
        return 2.0 * k1 * n * s + 3.0 * n * s + 2.0 * n * s
    return 3.0 * k1 * n * s + 4.0 * n * s + 2.0 * n * s      # fused CGS2: 3 passes


def summarise(table, n):
    """cycle averages and fixed-width numbers from {(orth, dtype, m): {k: ms}}"""
    out = {}
    for (orth, dt, m), per_k in sorted(table.items()):
        s = 4 if dt == "f32" else 8
        ks = sorted(per_k)
        avg = sum(per_k[k] for k in ks) / len(ks)
        avg_bytes = sum(bytes_add_vector(orth, k + 1, n, s) for k in ks) / len(ks)
        last = per_k[ks[-1]] if ks[-1] == m - 1 else None
        out[f"{orth}_{dt}_m{m}"] = {"cycle_avg_ms": round(avg, 4), "cycle_avg_GBps": round(avg_bytes / (avg * 1e-3) / 1e9, 1),
                                     "k_sampled": len(ks),
                                     "fixed_width_ms": None if last is None else round(last, 4),
                                     "fixed_width_GBps": None if last is None else round(bytes_add_vector(orth, m, n, s) / (last * 1e-3) / 1e9, 1)}
    return out


def gpu_sweep(ctx, g, spec, trials, dev):
    import numpy as np
    import torch
    rm, ind, val = ctx.gen(spec)
    n, nnz = rm.numel() - 1, ind.numel()
    A = g.CSR(ctx, rm, ind)

    def timed(fn, reset=None):
        if reset:
            reset()
        fn(); torch.cuda.synchronize()      # warm pass
        ts = []
        for _ in range(trials):
            if reset:
                reset()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]

    res = {"matrix": spec, "n": n, "nnz": nnz, "spmv": {}, "launches": 0}
    l0 = ctx.launches()
    for dt, tdt in (("f32", torch.float32), ("f64", torch.float64)):
        v = val.to(tdt)
        x = torch.from_numpy(ctx.rand_vect(n, 42)).to(dev).to(tdt)
        y = torch.empty_like(x)
        s = 4 if dt == "f32" else 8
        by = nnz * (s + 4) + 4 * (n + 1) + 2 * n * s
        ms = timed(lambda: ctx.spmv(A, v, 1.0, x, 0.0, y))
        res["spmv"][f"csr_{dt}"] = {"ms": round(ms, 4), "GBps": round(by / (ms * 1e-3) / 1e9, 1)}
        P = g.Packed(ctx, A, v)
        if P:
            ms = timed(lambda: ctx.spmv_packed(P, 1.0, x, 0.0, y))
            res["spmv"][f"packed_{dt}"] = {"ms": round(ms, 4), "GBps": round(by / (ms * 1e-3) / 1e9, 1)}
        del P, v, x, y
    table = {}
    for dt, tdt in (("f32", torch.float32), ("f64", torch.float64)):
        m = max(MS)
        ldv = (n + 31) // 32 * 32
        gen = torch.Generator(device=dev); gen.manual_seed(45)
        V = (torch.rand(ldv * (m + 1), dtype=tdt, device=dev, generator=gen) / (n ** 0.5))
        w0 = torch.rand(n, dtype=tdt, device=dev, generator=gen)
        w = torch.empty_like(w0)
        hcol = torch.zeros(m + 2, dtype=tdt, device=dev)
        for orth in ORTHS:
            for mm in MS:
                per_k = {}
                for k in ks_for(mm, True):
                    per_k[k] = timed(lambda: ctx.add_vector(orth, n, k, V, ldv, w, hcol), reset=lambda: w.copy_(w0))
                table[(orth, dt, mm)] = per_k
        del V, w0, w
        torch.cuda.empty_cache()
    res["add_vector"] = summarise(table, n)
    res["launches"] = int(ctx.launches() - l0)
    return res


def pair_rate(spmv_ms, cgs2_cycle_ms):
    return 1e3 / (spmv_ms + cgs2_cycle_ms)


def run_b200(args, rank, world, local_rank):
    if rank != 0:
        return
    sys.path.insert(0, ROOT)
    import torch
    import gmres_b200 as g
    from bench import ClockSampler, measured_peaks, force_host_threads
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    ctx = g.Context(local_rank)
    for kv in args.tune:
        ctx.set_tuning(kv.split("=")[0], int(kv.split("=")[1]))
    peak, peak_src = measured_peaks()
    trials = max(3, args.steps)
    sampler = ClockSampler(local_rank)
    t0 = time.perf_counter()
    small = gpu_sweep(ctx, g, args.sweep_matrix, trials, dev)
    big = gpu_sweep(ctx, g, "cd27:256", trials, dev) if args.sweep_matrix != "cd27:256" else small
    clocks = sampler.stop()
    wall = time.perf_counter() - t0
    sp = small["spmv"].get("packed_f32", small["spmv"]["csr_f32"])
    value = pair_rate(sp["ms"], small["add_vector"]["cgsr_f32_m100"]["cycle_avg_ms"])
    cpu = None
    if not args.no_cpu_baseline:
        force_host_threads()
        c = cpu_sweep(args.sweep_matrix)
        cpu = {"value": c["value"], "unit": UNIT, "cores": c["cores"], "kind": c["kind"], "sample": c["sample"], "sweep": c["sweep"]}
    dom = big["add_vector"]["cgsr_f32_m100"]
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": trials, "warmup": 1, "ms_per_step": 1e3 / value,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 and f64 (both swept)", "data": "synthetic",
            "config": {"workload": "sweep", "matrix": args.sweep_matrix, "also": "cd27:256", "restart_lengths": list(MS), "orths": list(ORTHS),
                       "timing": f"CUDA events per call, warm pass + median of {trials}; every operand >> 126 MB L2 except the narrowest bases",
                       "wall_s": round(wall, 1)},
            "roofline": {"bound": "hbm", "kernel": "add_vector CGS2 fp32, m = 100 cycle average, cd27:256", "achieved": dom["cycle_avg_GBps"], "peak": peak,
                         "unit": "GB/s", "frac": round(dom["cycle_avg_GBps"] / peak, 4), "traffic": None, "peak_source": peak_src},
            "sweep": {"b200_" + args.sweep_matrix: small, "b200_cd27:256": big}, "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                    "note": "per-kernel sweep: operands are device-resident by definition (kernel_perf_test.cpp times kernels, not transfers)"},
            "gpu_launches": small["launches"] + (big["launches"] if big is not small else 0), "clocks": clocks}
    print(json.dumps(line), flush=True)


def cpu_sweep(spec):
    """the reference's own MKL kernels (oracle/_ref) on `spec`; returns dict(value, cores, kind, sample, sweep)"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle as orc
    import oracle_ref
    import torch
    from bench import host_threads
    torch.set_num_threads(host_threads())
    if not oracle_ref.available():
        raise SystemExit("sweep: oracle/_ref is not built (needs /root/reference at build time)")
    rm, ind, val = orc.gen(spec)
    n = len(rm) - 1
    t0 = time.perf_counter()
    out = {"matrix": spec, "n": n, "nnz": int(len(ind)), "spmv": {}}
    for dt, isf in (("f32", True), ("f64", False)):
        ts = np.sort(oracle_ref.sweep_spmv(rm, ind, val, isf, 3))
        s = 4 if isf else 8
        by = len(ind) * (s + 4) + 4 * (n + 1) + 2 * n * s
        out["spmv"][f"mkl_{dt}"] = {"ms": round(float(ts[1]) * 1e3, 3), "GBps": round(by / float(ts[1]) / 1e9, 1)}
    table = {}
    for dt, isf in (("f32", True), ("f64", False)):
        for orth in ORTHS:
            for m in MS:
                ks = ks_for(m, False)
                if ks[-1] != m - 1:
                    ks.append(m - 1)
                sec = np.sort(oracle_ref.sweep_add_vector(n, m, orth, isf, ks, 3), axis=1)[:, 1]
                table[(orth, dt, m)] = {k: float(v) * 1e3 for k, v in zip(ks, sec)}
    out["add_vector"] = summarise(table, n)
    value = pair_rate(out["spmv"]["mkl_f32"]["ms"], out["add_vector"]["cgsr_f32_m100"]["cycle_avg_ms"])
    return {"value": value, "cores": int(oracle_ref.num_threads()), "kind": "reference",
            "sample": (f"{spec}: the reference's kernels_mkl.cpp spmv + Orthogonalization.hpp add_vector (MGS / CGS / CGS2, fp32 and fp64) built in oracle/_ref, "
                       f"warm pass + median of 3, cycle averages over a uniform subsample of k; {time.perf_counter() - t0:.1f} s of CPU work"),
            "sweep": out}


def run_reference(args):
    c = cpu_sweep(args.sweep_matrix)
    assert c["cores"] > 1 or (os.cpu_count() or 1) == 1
    line = {"impl": "reference", "metric": METRIC, "value": c["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": 3, "warmup": 1,
            "steps_requested": args.steps, "warmup_requested": args.warmup, "ms_per_step": 1e3 / c["value"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32 and f64 (both swept)", "data": "synthetic",
            "config": {"workload": "sweep", "matrix": args.sweep_matrix, "restart_lengths": list(MS), "orths": list(ORTHS)},
            "sweep": {"mkl_" + args.sweep_matrix: c["sweep"]},
            "cpu_baseline": {"value": c["value"], "unit": UNIT, "cores": c["cores"], "kind": c["kind"], "sample": c["sample"]},
            "e2e": {"value": c["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)
