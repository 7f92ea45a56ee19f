"""SpMV kernel variants on the bench matrices: achieved GB/s (algorithmic bytes / CUDA-event time)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gmres_b200 as g
ctx = g.Context(0)
for spec in (sys.argv[1:] or ["cd27:256", "lap2d:2048", "powerlaw:8000000"]):
    rm, ind, val = ctx.gen(spec)
    A = g.CSR(ctx, rm, ind)
    n = rm.numel() - 1
    x32 = torch.randn(n, dtype=torch.float32, device="cuda:0"); y32 = torch.empty_like(x32)
    x64 = x32.double(); y64 = torch.empty_like(x64); b64 = torch.randn_like(x64)
    v32 = val.float()
    ref = {}
    for variant in (0, 1):
        ctx.set_tuning("spmv_variant", variant)
        for name, cls, fn, out in [("spmv_f32", "spmv_f32", lambda: ctx.spmv(A, v32, 1.0, x32, 0.0, y32), y32),
                                   ("spmv_f64", "spmv_f64", lambda: ctx.spmv(A, val, 1.0, x64, 0.0, y64), y64),
                                   ("residual_f64_cast", "spmv_f64", lambda: ctx.residual_cast(A, val, b64, x64, None, y32), y32)]:
            fn()
            ctx.prof_enable(True); ctx.prof_reset()
            for _ in range(10):
                fn()
            p = ctx.prof_get(); ctx.prof_enable(False)
            o = out.double().clone()
            if variant == 0:
                ref[name] = o
                dev = 0.0
            else:
                dev = float((o - ref[name]).abs().max() / ref[name].abs().max())
            print(f"{spec:18s} variant {variant} {name:18s} {p[cls]['ms'] / 10:8.3f} ms {p[cls]['bytes'] / p[cls]['ms'] / 1e6:8.0f} GB/s   max rel diff vs variant 0: {dev:.2e}")
    del rm, ind, val, A, v32
    torch.cuda.empty_cache()
