#!/usr/bin/env python
"""gmres_perf_test — the reference's command-line harness (gmres_perf_test.cpp:309-455) over the B200 backend's
device-resident drivers.  Same flags, same stdout lines (the ones automated.py:33-38 scrapes), plus --gen for the
synthetic matrices (the reference only reads MatrixMarket files) and --json.

  python gmres_perf_test.py --Apath A.mtx --mode mixed --orth cgsr --prec identity --rlen 100 [--gpu]
  python gmres_perf_test.py --gen cd27:256 --rlen 100 --orth cgsr --prec identity

Flags as the reference: --Apath --bpath --rlen --rtol --repeat-iter --orthloss --tol --max-restarts --rand
--mode {mixed,baseline,single-prec,single} --orth {cgs,mgs,cgsr} --prec {identity,jacobi,ilu_jacobi} --jacobi-steps --gpu
(accepted, always on: there is no CPU path).  Defaults follow gmres_perf_test.cpp:313-325 except --prec (the exact-ILU
triangular solves are not provided -> identity).  --csv FILE appends the row automated.py:158-168 writes to
history-<matrix>.csv (matrix, mode key, orth, rlen, rtol, rorth, tol, cuda, prec, i, total_iters, res, err, ilu s, gmres s)."""
import csv
import json
import os
import sys
import time

import numpy as np


def main(argv):
    a = dict(Apath=None, bpath=None, gen=None, rlen=0, rtol=0.0, orthloss=False, repeat=False, tol=1e-6, max_restarts=1000000, rand=42,
             orth="mgs", mode="mixed", prec="identity", json=False, jacobi_steps=1, csv=None)
    i = 1
    while i < len(argv):
        f = argv[i]
        def val():
            nonlocal i
            i += 1
            return argv[i]
        if f == "--Apath": a["Apath"] = val()
        elif f == "--bpath": a["bpath"] = val()
        elif f == "--gen": a["gen"] = val()
        elif f == "--rlen": a["rlen"] = int(val())
        elif f == "--rtol": a["rtol"] = float(val())
        elif f == "--repeat-iter": a["repeat"] = True
        elif f == "--orthloss": a["orthloss"] = True
        elif f == "--tol": a["tol"] = float(val())
        elif f == "--max-restarts": a["max_restarts"] = int(val())
        elif f == "--rand": a["rand"] = int(val())
        elif f == "--mode":
            a["mode"] = val()
            if a["mode"] not in ("mixed", "baseline", "single-prec", "single"):
                print("Unknown test mode"); return 1
        elif f == "--orth":
            a["orth"] = val()
            if a["orth"] not in ("cgs", "mgs", "cgsr"):
                print("Unknown Orthogonalization"); return 1
        elif f == "--prec":
            a["prec"] = val()
            if a["prec"] == "ilu":
                print("Preconditioner ilu (exact triangular solves) is not provided by the B200 backend (identity, jacobi, ilu_jacobi)"); return 1
            if a["prec"] not in ("identity", "jacobi", "ilu_jacobi"):
                print("Unknown Preconditioner"); return 1
        elif f == "--jacobi-steps": a["jacobi_steps"] = int(val())
        elif f == "--csv": a["csv"] = val()
        elif f == "--gpu": pass
        elif f == "--json": a["json"] = True
        else:
            print("Unknown flag" + f); return 1          # gmres_perf_test.cpp:390-393
        i += 1
    if a["repeat"] and a["orthloss"]:
        print("Repeated Iteration Restart cannot be used with OrthLoss restart"); return 1
    if a["Apath"] is None and a["gen"] is None:
        print("No value suplied for A"); return 1       # :401-404 (sic)

    import torch
    import gmres_b200 as g
    ctx = g.Context(0)
    dev = "cuda:0"
    if a["gen"]:
        rm, ind, val = ctx.gen(a["gen"])
    else:
        rm_h, ind_h, val_h = g.read_matrix_market(a["Apath"])          # LoadMatrix<double>, :408
        rm, ind, val = (torch.from_numpy(t).to(dev) for t in (rm_h, ind_h, val_h))
    n = rm.numel() - 1
    A = g.CSR(ctx, rm, ind)
    if a["bpath"] is None:
        xt = torch.from_numpy(ctx.rand_vect(n, a["rand"])).to(dev)      # rand_vect + b = A x, :413-416
        b = torch.zeros(n, dtype=torch.float64, device=dev)
        ctx.spmv(A, val, 1.0, xt, 0.0, b)
    else:
        xt = torch.zeros(n, dtype=torch.float64, device=dev)            # x_host = 0, b = LoadVector(bpath), :417-421
        b_h = g.read_matrix_market_vector(a["bpath"])
        if len(b_h) != n:
            print(f"right-hand side has {len(b_h)} rows, the matrix {n}"); return 1
        b = torch.from_numpy(b_h).to(dev)
    print(f"||x|| = {ctx.nrm2(xt):g}")                                   # :223-225
    print(f"||b|| = {ctx.nrm2(b):g}")
    print(f"||A|| = {ctx.nrm2(val):g}")
    print("Doing Mixed Precision test" if a["mode"] == "mixed" else "Doing Baseline test")   # :61,128
    # alloc_convergence, :185-196
    conv = "base" if a["rtol"] == 0 else ("repeat" if a["repeat"] else ("orthloss" if a["orthloss"] else "relprecres"))
    t0 = time.perf_counter()
    val32 = torch.empty(val.numel(), dtype=torch.float32, device=dev)
    ctx.copy(val, val32)                                                 # SparseMatrix<float>(A): the "ilu took" window, :135-163
    ctx.sync()
    prec_s = time.perf_counter() - t0
    x = torch.zeros(n, dtype=torch.float64, device=dev)
    t0 = time.perf_counter()
    r = ctx.gmres(A, val, b, x, vals32=val32, mode=a["mode"], orth=a["orth"], conv=conv, prec=a["prec"], rlen=a["rlen"], tol=a["tol"],
                  rtol=a["rtol"], max_restarts=a["max_restarts"], jacobi_steps=a["jacobi_steps"], hist_cap=1)
    ctx.sync()
    gmres_s = time.perf_counter() - t0                                   # wall clock around the solver call, :165-167
    if r["status"] == 1:
        print(f"Found solution with rel prec res norm = {r['rel_prec_res']:g} when k = 0 and i = {r['outer_i']}")   # gmres.cpp:186-187
        print(f"  total iterations = {r['total_iters']}")
    else:
        print(f"Aborting after {r['total_iters']} iterations")           # gmres.cpp:190
    res = b.clone(); ctx.spmv(A, val, -1.0, x, 1.0, res)                 # :169-175
    res_norm, err_norm = ctx.nrm2(res), ctx.nrm2(x - xt)
    print(f"  ilu took {prec_s:g}s; gmres took {gmres_s:g}s")            # :177-178
    print(f"  resNorm = {res_norm:g}; errNorm = {err_norm:g}")
    if a["csv"]:
        mat = a["gen"] or os.path.splitext(os.path.basename(a["Apath"]))[0]
        key = {"baseline": "b", "mixed": "mp", "single-prec": "p", "single": "s"}[a["mode"]]
        rt = ("R" if a["repeat"] else "") + (f"{a['rtol']:g}" if not a["orthloss"] else "0")
        ro = f"{a['rtol']:g}" if a["orthloss"] else "0"
        with open(a["csv"], "a", newline="") as fcsv:
            csv.writer(fcsv, delimiter=",").writerow([mat, key, a["orth"].upper(), a["rlen"], rt, ro, f"{a['tol']:g}", "cuda", a["prec"], r["outer_i"],
                                                      r["total_iters"], f"{res_norm:g}", f"{err_norm:g}", f"{prec_s:g}", f"{gmres_s:g}"])
    if a["json"]:
        print(json.dumps(dict(n=n, nnz=int(ind.numel()), status=r["status"], i=r["outer_i"], total_iterations=r["total_iters"],
                              restarts=r["total_restarts"], gmres_s=gmres_s, solve_ms=r["solve_ms"], resNorm=res_norm, errNorm=err_norm)))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
