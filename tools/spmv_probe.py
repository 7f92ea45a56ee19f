"""SpMV probe for ncu captures and quick timings: generates a matrix on the device and runs the CSR kernel (fp32, fp64), the
fused fp64 residual and - when the structure packs - the packed kernels a few times, printing CUDA-event times.
    python tools/spmv_probe.py --gen powerlaw:8000000 [--reps 5] [--tune key=val ...]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gen", default="powerlaw:8000000")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--tune", action="append", default=[])
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    import torch
    import gmres_b200 as g
    ctx = g.Context(0)
    for kv in args.tune:
        ctx.set_tuning(kv.split("=")[0], int(kv.split("=")[1]))
    rm, ind, val = ctx.gen(args.gen)
    n, nnz = rm.numel() - 1, ind.numel()
    A = g.CSR(ctx, rm, ind)
    v32 = val.float()
    x32 = torch.rand(n, dtype=torch.float32, device="cuda:0")
    x64 = x32.double()
    b64 = torch.rand(n, dtype=torch.float64, device="cuda:0")
    y32 = torch.empty_like(x32); y64 = torch.empty_like(x64); w32 = torch.empty_like(x32)
    out = {"matrix": args.gen, "n": n, "nnz": nnz, "ms": {}, "GBps": {}}
    bytes32 = nnz * 8 + 4 * (n + 1) + 2 * n * 4
    bytes64 = nnz * 12 + 4 * (n + 1) + 2 * n * 8
    bytes_res = nnz * 12 + 4 * (n + 1) + 8 * n + 8 * n + 4 * n
    P32 = g.Packed(ctx, A, v32)
    P64 = g.Packed(ctx, A, val)
    runs = {"csr_f32": (lambda: ctx.spmv(A, v32, 1.0, x32, 0.0, y32), bytes32),
            "csr_f64": (lambda: ctx.spmv(A, val, 1.0, x64, 0.0, y64), bytes64),
            "residual_f64": (lambda: ctx.residual_cast(A, val, b64, x64, None, w32), bytes_res)}
    if P32:
        runs["packed_f32"] = (lambda: ctx.spmv_packed(P32, 1.0, x32, 0.0, y32), bytes32)
    if P64:
        runs["packed_f64"] = (lambda: ctx.spmv_packed(P64, 1.0, x64, 0.0, y64), bytes64)
    for name, (fn, nbytes) in runs.items():
        if args.only and name not in args.only.split(","):
            continue
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        out["ms"][name] = round(ts[len(ts) // 2], 4)
        out["GBps"][name] = round(nbytes / (ts[len(ts) // 2] * 1e-3) / 1e9, 1)
    if P32:
        G, off, _, _ = (None, None, None, None)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
