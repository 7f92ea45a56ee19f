"""Kernel-level sweep on the GPU: per-pass achieved GB/s of the orthogonalisation kernels vs basis width, for the
kernel variants selectable through mpg_set_tuning, plus SpMV fp32/fp64 on the bench matrix.  Prints a table.
    python tools/tune.py [--n 16777216] [--spec cd27:256]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gmres_b200 as g

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=16777216)
ap.add_argument("--m", type=int, default=100)
ap.add_argument("--spec", default="cd27:256")
ap.add_argument("--ks", default="1,2,4,8,16,25,32,50,64,75,100")
ap.add_argument("--variants", default="default,stages2,stages1,noserp,passA_rb,rb16k")
ap.add_argument("--no-spmv", action="store_true")
ap.add_argument("--fuse-min", type=int, default=1, help="library default of fuse_min_cols (restored between variants)")
args = ap.parse_args()
FUSE_MIN_DEFAULT = args.fuse_min

ctx = g.Context(0)
n, m = args.n, args.m
ldv = (n + 31) // 32 * 32
V = torch.empty(ldv * (m + 1), dtype=torch.float32, device="cuda:0")
for j in range(m + 1):
    V[j * ldv:(j + 1) * ldv].normal_()
V *= (1.0 / n ** 0.5)
w0 = torch.randn(ldv + 32, dtype=torch.float32, device="cuda:0")
w = torch.empty_like(w0)
h = torch.zeros(m + 2, dtype=torch.float32, device="cuda:0")

VARIANTS = {
    "default": {},
    "stages1": {"vpass_stages": 1},
    "stages2": {"vpass_stages": 2},
    "noserp": {"vpass_serpentine": 0},
    "passA_rb": {"passA_rb": 1},
    "unfused": {"cgs2_fused": 0},
    "rb4k": {"passA_rb": 1, "gemvt_rows_per_block": 4096},
    "rb16k": {"passA_rb": 1, "gemvt_rows_per_block": 16384},
    "direct16": {"vdirect_max_cols_a": 16, "vdirect_max_cols_b": 16},
    "vrow48": {"vrow_max_cols": 48, "vrow_max_cols_a": 48},
    "vrow64": {"vrow_max_cols": 64, "vrow_max_cols_a": 64},
    "vrow_all": {"vdirect_max_cols_a": 0, "vdirect_max_cols_b": 0, "vrow_max_cols": 32},   # TMA row-owner kernel for every k1 <= 32
    "vpass_all": {"vdirect_max_cols_a": 0, "vdirect_max_cols_b": 0, "vrow_max_cols": 0},   # TMA column-owner kernel for every k1
    "regs_lt16": {"vdirect_max_cols_a": 0, "vdirect_max_cols_b": 0, "fuse_min_cols": 16, "vrow_max_cols": 0},   # round-1a policy
}
DEFAULTS = {"vpass_stages": 0, "vpass_serpentine": 1, "passA_rb": 0, "cgs2_fused": 1,
            "gemvt_rows_per_block": 8192, "fuse_min_cols": FUSE_MIN_DEFAULT, "vrow_max_cols": 56,
            "vdirect_max_cols_a": 8, "vdirect_max_cols_b": 8, "vrow_max_cols_a": 32}


def run(orth, k, reps=3):
    out = {}
    for it in range(reps + 1):
        w.copy_(w0)
        if it == 1:
            ctx.prof_enable(True); ctx.prof_reset()
        ctx.add_vector(orth, n, k - 1, V, ldv, w, h)
    p = ctx.prof_get()
    ctx.prof_enable(False)
    for c in ("vpass", "gemvn", "gemvt", "elementwise"):
        if p[c]["launches"]:
            out[c] = (p[c]["ms"] / reps, p[c]["bytes"] / reps)
    return out


def gbs(t):
    return t[1] / (t[0] * 1e-3) / 1e9 if t and t[0] > 0 else 0.0


print(f"n={n} ldv={ldv} fp32; GB/s = algorithmic bytes / CUDA-event time; A = h=V'w, B = fused w-=Vh,c=V'w, C = w-=Vc,norm")
for vname in args.variants.split(","):
    for k_, v_ in DEFAULTS.items():
        ctx.set_tuning(k_, v_)
    for k_, v_ in VARIANTS[vname].items():
        ctx.set_tuning(k_, v_)
    print(f"--- variant {vname} {VARIANTS[vname]}")
    print(f"{'k1':>4} {'A ms':>8} {'A GB/s':>8} {'B ms':>8} {'B GB/s':>8} {'C ms':>8} {'C GB/s':>8} {'cgs2 ms':>8} {'cgs2 GB/s(3-pass bytes)':>10}")
    tot_ms = tot_b = 0.0
    for k1 in [int(x) for x in args.ks.split(",")]:
        a = run("cgs", k1)
        b = run("cgsr", k1)
        bytes3 = (3.0 * k1 + 4.0) * n * 4
        if vname in ("unfused",) or "vpass" not in b or "vpass" not in a and not (vname.startswith("passA_rb") or vname.startswith("rb")):
            ms = sum(x[0] for x in b.values() if x) - b.get("elementwise", (0, 0))[0]
            print(f"{k1:>4} {'':>8} {'':>8} {'':>8} {'':>8} {'':>8} {'':>8} {ms:8.3f} {bytes3 / ms / 1e6:10.1f}   gemvt {gbs(b.get('gemvt')):.0f} gemvn {gbs(b.get('gemvn')):.0f}")
            tot_ms += ms; tot_b += bytes3
            continue
        if vname.startswith("passA_rb") or vname.startswith("rb"):
            A = a.get("gemvt"); Bv = b.get("vpass")
        else:
            A = a.get("vpass")
            bv = b.get("vpass")
            Bv = (bv[0] - A[0], bv[1] - A[1]) if bv and A else None
        Cc = b.get("gemvn")
        ms = (A[0] if A else 0) + (Bv[0] if Bv else 0) + (Cc[0] if Cc else 0)
        print(f"{k1:>4} {A[0]:8.3f} {gbs(A):8.0f} {Bv[0]:8.3f} {gbs(Bv):8.0f} {Cc[0]:8.3f} {gbs(Cc):8.0f} {ms:8.3f} {bytes3 / ms / 1e6:10.1f}")
        tot_ms += ms; tot_b += bytes3
    if tot_ms:
        print(f"sum over listed k1: {tot_ms:.2f} ms, {tot_b / tot_ms / 1e6:.0f} GB/s")

if not args.no_spmv:
    for k_, v_ in DEFAULTS.items():
        ctx.set_tuning(k_, v_)
    rm, ind, val = ctx.gen(args.spec)
    A = g.CSR(ctx, rm, ind)
    nn = rm.numel() - 1
    x32 = torch.randn(nn, dtype=torch.float32, device="cuda:0"); y32 = torch.empty_like(x32)
    x64 = x32.double(); y64 = torch.empty_like(x64); b64 = torch.randn_like(x64)
    v32 = val.float()
    P32 = g.Packed(ctx, A, v32)
    P64 = g.Packed(ctx, A, val)
    cases = [("spmv_f32", lambda: ctx.spmv(A, v32, 1.0, x32, 0.0, y32)), ("spmv_f64", lambda: ctx.spmv(A, val, 1.0, x64, 0.0, y64)),
             ("residual_f64_cast", lambda: ctx.residual_cast(A, val, b64, x64, None, y32))]
    if P32:
        cases.append(("spmv_f32_packed", lambda: ctx.spmv_packed(P32, 1.0, x32, 0.0, y32)))
        cases.append(("pack_update_f32", lambda: P32.update(v32)))
    if P64:
        cases.append(("spmv_f64_packed", lambda: ctx.spmv_packed(P64, 1.0, x64, 0.0, y64)))
    for name, fn in cases:
        fn()
        ctx.prof_enable(True); ctx.prof_reset()
        for _ in range(5):
            fn()
        p = ctx.prof_get(); ctx.prof_enable(False)
        c = "elementwise" if name.startswith("pack_update") else ("spmv_f32" if "f32" in name else "spmv_f64")
        print(f"{name}: {p[c]['ms'] / 5:.3f} ms  {p[c]['bytes'] / p[c]['ms'] / 1e6:.0f} GB/s (algorithmic)")
