"""Per-CTA phase timestamps of one vpass launch (development aid): where the fixed cost of a V pass on a small slab goes.
    python tools/vpass_timeline.py [--n 2097152] [--k1 100]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import gmres_b200 as g

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=2097152)
ap.add_argument("--k1", default="100,64")
args = ap.parse_args()
ctx = g.Context(0)
n = args.n
ldv = (n + 31) // 32 * 32
m = 101
V = torch.randn(ldv * (m + 1), dtype=torch.float32, device="cuda:0") * (1.0 / n ** 0.5)
w0 = torch.randn(ldv + 32, dtype=torch.float32, device="cuda:0")
w = torch.empty_like(w0)
h = torch.zeros(m + 2, dtype=torch.float32, device="cuda:0")
dbg = torch.zeros(8 * 160, dtype=torch.int64, device="cuda:0")
names = ["start", "first tile in", "last tile done", "partials written", "ticket taken", "finish done (last CTA)"]
for k1 in [int(x) for x in args.k1.split(",")]:
    for orth, label in (("cgs", "pass A (h = V'w) + gemv-N"), ("cgsr", "pass A, pass B, gemv-N")):
        for rep in range(3):
            w.copy_(w0)
            ctx.debug_timing(dbg if rep == 2 else None)
            dbg.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ctx.add_vector(orth, n, k1 - 1, V, ldv, w, h)
            e1.record(); torch.cuda.synchronize()
        ctx.debug_timing(None)
        t = dbg.cpu().numpy().reshape(160, 8).astype(np.int64)
        t = t[t[:, 0] > 0]
        t0 = t[:, 0].min()
        print(f"n={n} k1={k1} {label}: add_vector {e0.elapsed_time(e1) * 1e3:.1f} us; timestamps of the LAST staged launch, {len(t)} CTAs, us after the first CTA start")
        for i, nm in enumerate(names):
            col = t[:, i]
            col = col[col > 0]
            if len(col):
                print(f"   {nm:26s} min {(col.min() - t0) / 1e3:8.2f}  median {(np.median(col) - t0) / 1e3:8.2f}  max {(col.max() - t0) / 1e3:8.2f}")
