"""Generate tests/golden/gmres_cases.json from the REFERENCE ITSELF: oracle/_ref/libref.so is the reference's own
gmres.cpp / Orthogonalization.hpp / IterUtil.hpp / kernels_mkl.cpp, compiled unmodified by oracle/ref.mk against the
oneMKL inside libtorch_cpu.so.  Run in the container that has /root/reference:

    make -C oracle ref && python tests/golden/make_goldens.py

For every case it stores what the reference produced (status, restart / iteration counts, residual histories, the
printed resNorm / errNorm) and how far the oracle's history is from it ("dev_oracle_vs_ref": max relative difference
while the residual is above 1e-4 of its starting value).  That deviation between two correct implementations of the
same algorithm is the yardstick for the GPU parity envelope (tests/test_solver_gpu.py)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as orc  # noqa: E402
import oracle_ref as ref  # noqa: E402

CASES = []
for orth in ("cgsr", "cgs", "mgs"):
    CASES += [dict(spec="lap2d:64", mode="mixed", orth=orth, rlen=50, tol=1e-9),
              dict(spec="cd27:16", mode="mixed", orth=orth, rlen=100, tol=1e-9),
              dict(spec="powerlaw:5000", mode="mixed", orth=orth, rlen=50, tol=1e-9)]
CASES += [dict(spec="lap2d:7", mode="mixed", orth="cgsr", rlen=10, tol=1e-9),
          dict(spec="cd27:3", mode="mixed", orth="cgsr", rlen=5, tol=1e-9)]
for mode in ("baseline", "single-prec", "single"):
    tol = 1e-10 if mode == "baseline" else 1e-6
    CASES += [dict(spec="lap2d:48", mode=mode, orth="cgsr", rlen=50, tol=tol), dict(spec="cd27:12", mode=mode, orth="cgsr", rlen=30, tol=tol)]
CASES += [dict(spec="lap2d:40", mode="mixed", orth="cgsr", rlen=40, tol=1e-9, conv="relprecres", rtol=1e-2),
          dict(spec="lap2d:40", mode="mixed", orth="cgsr", rlen=40, tol=1e-9, conv="repeat", rtol=1e-2),
          dict(spec="lap2d:40", mode="mixed", orth="cgsr", rlen=40, tol=1e-9, conv="orthloss", rtol=1e-3)]
for mode in ("mixed", "baseline", "single-prec"):
    CASES += [dict(spec="powerlaw:4000", mode=mode, orth="cgsr", rlen=20, tol=1e-9 if mode != "single-prec" else 1e-6, prec="jacobi")]
# BASELINE.json configs[0]: 2-D Laplacian 512^2, fp64 GMRES(50) CGS2 on the MKL CPU backend (reference gmres_perf_test path)
CASES += [dict(spec="lap2d:512", mode="baseline", orth="cgsr", rlen=50, tol=1e-6),
          dict(spec="lap2d:512", mode="mixed", orth="cgsr", rlen=50, tol=1e-6)]
# restart policies (IterUtil.hpp:84-227) on the nonsymmetric stencil and on power-law rows as well
for spec, rl in (("cd27:14", 40), ("powerlaw:3000", 30)):
    CASES += [dict(spec=spec, mode="mixed", orth="cgsr", rlen=rl, tol=1e-9, conv="relprecres", rtol=1e-2),
              dict(spec=spec, mode="mixed", orth="cgsr", rlen=rl, tol=1e-9, conv="repeat", rtol=1e-2)]
    if not spec.startswith("powerlaw"):   # (LostOrthogonality_ on power-law rows: ~3300 one-step cycles decided on fp32 noise; even two
        #                                    runs of the reference itself differ there, so it cannot serve as a fixture)
        CASES += [dict(spec=spec, mode="mixed", orth="cgsr", rlen=rl, tol=1e-9, conv="orthloss", rtol=1e-3)]
# right-hand sides scaled far outside the range where an UNSCALED fp32 sum of squares works (|b|^2 overflows / underflows in
# fp32): the reference's BLAS nrm2 is scaled (kernels_mkl.cpp:97-115), so these pin the nrm2 semantics end to end
for bs in (1e25, 1e-25):
    CASES += [dict(spec="cd27:12", mode="mixed", orth="cgsr", rlen=30, tol=1e-9, bscale=bs),
              dict(spec="lap2d:40", mode="single", orth="cgsr", rlen=40, tol=1e-6, bscale=bs)]

HIST_KEEP = 600
FLOOR = 1e-4


def deviation(h, h0):
    m = min(len(h), len(h0))
    if m == 0:
        return 0.0
    a, b = np.asarray(h[:m]), np.asarray(h0[:m])
    live = b >= FLOOR * b[0]
    return float((np.abs(a - b) / np.maximum(b, 1e-300))[live].max()) if live.any() else 0.0


def deviation_first_cycle(h, h0, rlen):
    """the first restart cycle while the residual is above 1e-2 of its start: no restart has fed rounding noise back yet"""
    m = min(len(h), len(h0), rlen)
    if m == 0:
        return 0.0
    a, b = np.asarray(h[:m]), np.asarray(h0[:m])
    live = b >= 1e-2 * b[0]
    return float((np.abs(a - b) / np.maximum(b, 1e-300))[live].max()) if live.any() else 0.0


def main():
    assert ref.available(), "build oracle/_ref first (make -C oracle ref)"
    out = {"generator": "tests/golden/make_goldens.py", "reference": "oracle/_ref/libref.so (reference sources + oneMKL in libtorch_cpu.so)",
           "mkl_threads": ref.num_threads(), "hist_floor": FLOOR, "cases": []}
    for c in CASES:
        kw = {k: v for k, v in c.items() if k not in ("spec", "bscale")}
        rm, ind, val = orc.gen(c["spec"])
        n = len(rm) - 1
        xt = orc.rand_vect(n, 42)
        b = np.zeros(n)
        orc.spmv(rm, ind, val, 1.0, xt, 0.0, b)
        if "bscale" in c:
            xt = xt * c["bscale"]; b = b * c["bscale"]
        rr = ref.gmres(rm, ind, val, b, true_x=xt, max_restarts=5000, **kw)
        ro = orc.gmres(rm, ind, val, b, max_restarts=5000, **kw)
        res_o = b.copy(); orc.spmv(rm, ind, val, -1.0, ro["x"], 1.0, res_o)
        g = dict(c)
        g.update(n=n, nnz=int(len(val)),
                 ref=dict(status=rr["status"], total_iters=rr["total_iters"], total_restarts=rr["total_restarts"], outer_i=rr["outer_i"],
                          rel_prec_res=rr["rel_prec_res"], res_norm=rr["res_norm"], err_norm=rr["err_norm"],
                          hist_inner=[float(v) for v in rr["hist_inner"][:HIST_KEEP]],
                          hist_outer=[[float(v) for v in row] for row in rr["hist_outer"][:50]]),
                 oracle=dict(status=ro["status"], total_iters=ro["total_iters"], total_restarts=ro["total_restarts"], outer_i=ro["outer_i"],
                             res_norm=float(orc.nrm2(res_o)), err_norm=float(orc.nrm2(ro["x"] - xt))),
                 dev_oracle_vs_ref=deviation(ro["hist_inner"], rr["hist_inner"]),
                 dev_first_cycle=deviation_first_cycle(ro["hist_inner"], rr["hist_inner"], c["rlen"]))
        out["cases"].append(g)
        print(f"{c['spec']:14s} {c['mode']:11s} {c['orth']:4s} ref it={rr['total_iters']} rs={rr['total_restarts']} | oracle it={ro['total_iters']} "
              f"rs={ro['total_restarts']} | dev {g['dev_oracle_vs_ref']:.2e} | resNorm ref {rr['res_norm']:.3e} orc {g['oracle']['res_norm']:.3e}")
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "gmres_cases.json"), "w") as f:
        json.dump(out, f)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
