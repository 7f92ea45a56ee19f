"""BASELINE.json configs[4]: the kernel_perf_test sweep (the reference's kernel_perf_test.cpp no longer builds; its shape
is the spec: ops x {float,double}, one warm pass then timed passes, --vcols, basis/vectors from mt19937 floats).
SpMV fp32 vs fp64 on a matrix, and one Arnoldi step (add_vector = orthogonalise + norm + normalise) with MGS / CGS / CGS2
at fixed basis width k+1 in {25, 50, 100}; achieved GB/s = algorithmic bytes / CUDA-event time.
With --cpu the same ops are timed through the oracle's CPU restatement (OpenMP) on the host cores.
    python tools/kernel_perf_test.py [--gen cd27:128] [--vcols 25,50,100] [--cpu]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import gmres_b200 as g

ap = argparse.ArgumentParser()
ap.add_argument("--gen", default="cd27:256")
ap.add_argument("--vcols", default="25,50,100")
ap.add_argument("--trials", type=int, default=5)
ap.add_argument("--cpu", action="store_true")
ap.add_argument("--cpu-gen", default="cd27:64")
args = ap.parse_args()

ctx = g.Context(0)
rm, ind, val = ctx.gen(args.gen)
A = g.CSR(ctx, rm, ind)
n, nnz = rm.numel() - 1, ind.numel()
out = {"matrix": args.gen, "n": n, "nnz": nnz, "gpu": {}, "cpu": {}}


def timed(cls, fn, trials):
    fn()  # warm pass (kernel_perf_test.cpp:170-179)
    ctx.prof_enable(True); ctx.prof_reset()
    for _ in range(trials):
        fn()
    p = ctx.prof_get(); ctx.prof_enable(False)
    ms = sum(v["ms"] for k, v in p.items() if k in cls) / trials
    return ms


for dt, name in [(torch.float32, "float"), (torch.float64, "double")]:
    s = 4 if dt == torch.float32 else 8
    v = val.to(dt)
    x = torch.from_numpy(ctx.rand_vect(n, 42)).to("cuda:0").to(dt)
    y = torch.empty_like(x)
    ms = timed(("spmv_f32", "spmv_f64"), lambda: ctx.spmv(A, v, 1.0, x, 0.0, y), args.trials)
    by = nnz * (s + 4) + 4 * (n + 1) + 2 * n * s
    out["gpu"][f"spmv_{name}"] = {"ms": ms, "GBps": by / ms / 1e6}
    P = g.Packed(ctx, A, v)   # the operator the solver (and SparseMatrix<T,B200>) actually multiplies with
    if P:
        ms = timed(("spmv_f32", "spmv_f64"), lambda: ctx.spmv_packed(P, 1.0, x, 0.0, y), args.trials)
        out["gpu"][f"spmv_{name}_packed"] = {"ms": ms, "GBps": by / ms / 1e6}
    del P
    ms = timed(("reduce",), lambda: ctx.dot(x, y), args.trials)
    out["gpu"][f"dot_{name}"] = {"ms": ms, "GBps": 2 * n * s / ms / 1e6}
    for k1 in [int(c) for c in args.vcols.split(",")]:
        ldv = (n + 31) // 32 * 32
        V = torch.empty(ldv * (k1 + 1), dtype=dt, device="cuda:0")
        for j in range(k1 + 1):
            V[j * ldv:(j + 1) * ldv].normal_()
        V *= 1.0 / n ** 0.5
        w0 = torch.randn(n, dtype=dt, device="cuda:0"); w = torch.empty_like(w0)
        h = torch.zeros(k1 + 2, dtype=dt, device="cuda:0")
        for orth, passes in [("mgs", None), ("cgs", 2), ("cgsr", 3)]:
            def step():
                w.copy_(w0)
                ctx.add_vector(orth, n, k1 - 1, V, ldv, w, h)
            ms = timed(("vpass", "gemvn", "gemvt", "reduce", "elementwise"), step, args.trials) - 0.0
            # subtract the w.copy_ (torch kernel, not profiled by our classes): nothing to subtract
            by = (5 * k1 * n * s) if orth == "mgs" else ((2 * k1 + 3) * n * s if orth == "cgs" else (3 * k1 + 4) * n * s)
            by += 2 * n * s  # normalise
            out["gpu"][f"{orth}_{name}_k{k1}"] = {"ms": ms, "GBps": by / ms / 1e6}
        del V

if args.cpu:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    rmc, indc, valc = orc.gen(args.cpu_gen)
    nc = len(rmc) - 1
    out["cpu"]["matrix"] = args.cpu_gen
    out["cpu"]["threads"] = orc.num_threads()
    for dt, name in [(np.float32, "float"), (np.float64, "double")]:
        vv = valc.astype(dt); xx = orc.rand_vect(nc).astype(dt); yy = np.zeros(nc, dt)
        orc.spmv(rmc, indc, vv, 1.0, xx, 0.0, yy)
        t = time.perf_counter()
        for _ in range(args.trials):
            orc.spmv(rmc, indc, vv, 1.0, xx, 0.0, yy)
        ms = (time.perf_counter() - t) / args.trials * 1e3
        s = 4 if dt == np.float32 else 8
        out["cpu"][f"spmv_{name}"] = {"ms": ms, "GBps": (len(vv) * (s + 4) + 4 * (nc + 1) + 2 * nc * s) / ms / 1e6}
        for k1 in [int(c) for c in args.vcols.split(",")]:
            Vc = np.asfortranarray(np.random.default_rng(1).standard_normal((nc, k1 + 1)).astype(dt) / nc ** 0.5)
            for orth in ("mgs", "cgs", "cgsr"):
                ww = np.random.default_rng(2).standard_normal(nc).astype(dt)
                t = time.perf_counter()
                orc.add_vector(orth, Vc, k1 - 1, ww)
                ms = (time.perf_counter() - t) * 1e3
                by = (5 * k1 * nc * s) if orth == "mgs" else ((4 * k1 + 6) * nc * s if orth == "cgsr" else (2 * k1 + 3) * nc * s)
                out["cpu"][f"{orth}_{name}_k{k1}"] = {"ms": ms, "GBps": by / ms / 1e6}
print(json.dumps(out))
