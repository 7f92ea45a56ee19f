"""Drop-in proof: the reference's OWN command-line harness and solver driver (gmres_perf_test.cpp, gmres.cpp,
Orthogonalization.hpp, IterUtil.hpp, kernels.hpp, types.hpp — compiled unmodified from /root/reference by
`make -C oracle -f ref.mk b200`) linked against include/b200/{types_b200.hpp,kernels_b200.cpp} + libmpgmres_b200.so.
Its `--gpu` switch then runs every operator of the surface on the B200 backend; the same binary without `--gpu` runs
the reference's MKL path.  Both must report the same iteration counts and post-solve norms (stdout contract scraped
like automated.py:33-38).  The binary is built where /root/reference exists and travels to the GPU box."""
import os
import re
import subprocess

import numpy as np
import pytest

from util import problem

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "gmres_perf_test_b200")


def write_mtx(path, rm, ind, val):
    n = len(rm) - 1
    rows = np.repeat(np.arange(n), np.diff(rm))
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write(f"{n} {n} {len(val)}\n")
        for r_, c_, v_ in zip(rows, ind, val):
            f.write(f"{r_ + 1} {c_ + 1} {v_:.17g}\n")


def run_cli(mtx, gpu, mode, orth, prec, rlen, tol, extra=()):
    cmd = [EXE, "--Apath", str(mtx), "--mode", mode, "--orth", orth, "--prec", prec, "--rlen", str(rlen), "--tol", str(tol), *extra]
    if gpu:
        cmd.append("--gpu")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    m = re.search(r"Found solution with rel prec res norm = (\S+) when k = (\d+) and i = (\d+)\s+total iterations = (\d+)", out.stdout)
    assert m, out.stdout
    t = re.search(r"ilu took (\S+)s; gmres took (\S+)s", out.stdout)
    r = re.search(r"resNorm = (\S+); errNorm = (\S+)", out.stdout)
    return dict(i=int(m.group(3)), iters=int(m.group(4)), res=float(r.group(1)), err=float(r.group(2)), gmres_s=float(t.group(2)), stdout=out.stdout)


@pytest.mark.parametrize("spec,mode,orth,prec,rlen,tol", [
    ("lap2d:40", "mixed", "cgsr", "identity", 40, 1e-9),
    ("lap2d:40", "mixed", "mgs", "identity", 40, 1e-9),
    ("lap2d:40", "mixed", "cgs", "identity", 40, 1e-9),
    ("cd27:12", "mixed", "cgsr", "identity", 30, 1e-9),
    ("cd27:12", "baseline", "cgsr", "identity", 30, 1e-10),
    ("cd27:12", "single", "cgsr", "identity", 30, 1e-6),
    ("cd27:12", "single-prec", "cgsr", "jacobi", 30, 1e-6),
    ("powerlaw:3000", "mixed", "cgsr", "jacobi", 20, 1e-9),
])
def test_reference_cli_runs_on_b200_backend(orc, tmp_path, spec, mode, orth, prec, rlen, tol):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/gmres_perf_test_b200 not built (needs /root/reference: make -C oracle -f ref.mk b200)")
    rm, ind, val, xt, b = problem(orc, spec)
    mtx = tmp_path / "a.mtx"
    write_mtx(mtx, rm, ind, val)
    host = run_cli(mtx, False, mode, orth, prec, rlen, tol)
    gpu = run_cli(mtx, True, mode, orth, prec, rlen, tol)
    assert (gpu["i"], gpu["iters"]) == (host["i"], host["iters"]), (gpu["stdout"], host["stdout"])
    scale = tol * (np.linalg.norm(b) + np.linalg.norm(val.astype(np.float32)) * np.linalg.norm(xt))
    assert gpu["res"] <= max(8 * host["res"], scale)
    assert gpu["err"] <= max(8 * host["err"], 100 * tol * np.linalg.norm(xt))
    assert "Doing Mixed Precision test" in gpu["stdout"] or "Doing Baseline test" in gpu["stdout"]


def test_reference_cli_restart_policies_on_b200_backend(orc, tmp_path):
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/gmres_perf_test_b200 not built")
    rm, ind, val, xt, b = problem(orc, "lap2d:40")
    mtx = tmp_path / "a.mtx"
    write_mtx(mtx, rm, ind, val)
    for extra in (("--rtol", "1e-2"), ("--rtol", "1e-2", "--repeat-iter"), ("--rtol", "1e-3", "--orthloss")):
        host = run_cli(mtx, False, "mixed", "cgsr", "identity", 40, 1e-9, extra)
        gpu = run_cli(mtx, True, "mixed", "cgsr", "identity", 40, 1e-9, extra)
        assert abs(gpu["iters"] - host["iters"]) <= 0.05 * host["iters"] + 2, (extra, gpu["iters"], host["iters"])
        assert abs(gpu["i"] - host["i"]) <= 0.05 * host["i"] + 1


def test_python_harness_prints_the_reference_stdout_contract(orc, tmp_path):
    """gmres_perf_test.py (native device-resident driver) vs the reference CLI's MKL path on the same .mtx: same fields"""
    import sys
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/gmres_perf_test_b200 not built")
    rm, ind, val, xt, b = problem(orc, "cd27:10")
    mtx = tmp_path / "a.mtx"
    write_mtx(mtx, rm, ind, val)
    host = run_cli(mtx, False, "mixed", "cgsr", "identity", 30, 1e-9)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "gmres_perf_test.py"), "--Apath", str(mtx), "--mode", "mixed", "--orth", "cgsr",
                          "--prec", "identity", "--rlen", "30", "--tol", "1e-9", "--gpu"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    m = re.search(r"Found solution with rel prec res norm = (\S+) when k = (\d+) and i = (\d+)\s+total iterations = (\d+)", out.stdout)
    assert m and (int(m.group(3)), int(m.group(4))) == (host["i"], host["iters"]), out.stdout
    assert re.search(r"  ilu took \S+s; gmres took \S+s\n  resNorm = \S+; errNorm = \S+", out.stdout)
    assert out.stdout.startswith("||x|| = ")
    bad = subprocess.run([sys.executable, os.path.join(ROOT, "gmres_perf_test.py"), "--bogus"], capture_output=True, text=True)
    assert bad.returncode == 1 and "Unknown flag" in bad.stdout      # gmres_perf_test.cpp:390-393


def test_python_harness_bpath_and_csv_row(orc, tmp_path):
    """--bpath (LoadVector, gmres_perf_test.cpp:417-421: x_true = 0, errNorm degenerates to ||x||) and the history CSV row of
    automated.py:158-168, against the reference CLI's MKL path on the same matrix and right-hand-side files"""
    import csv
    import sys
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/gmres_perf_test_b200 not built")
    rm, ind, val, xt, b = problem(orc, "lap2d:30")
    mtx, rhs = tmp_path / "a.mtx", tmp_path / "b.mtx"
    write_mtx(mtx, rm, ind, val)
    with open(rhs, "w") as f:
        f.write(f"%%MatrixMarket matrix array real general\n{len(b)} 1\n")
        for v in b:
            f.write(f"{v:.17g}\n")
    host = run_cli(mtx, False, "mixed", "cgsr", "identity", 30, 1e-9, ("--bpath", str(rhs)))
    hist = tmp_path / "history-a.csv"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "gmres_perf_test.py"), "--Apath", str(mtx), "--bpath", str(rhs), "--mode", "mixed", "--orth", "cgsr",
                          "--prec", "identity", "--rlen", "30", "--tol", "1e-9", "--csv", str(hist), "--gpu"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("||x|| = 0\n")
    m = re.search(r"Found solution with rel prec res norm = (\S+) when k = (\d+) and i = (\d+)\s+total iterations = (\d+)", out.stdout)
    assert m and (int(m.group(3)), int(m.group(4))) == (host["i"], host["iters"]), out.stdout
    r = re.search(r"resNorm = (\S+); errNorm = (\S+)", out.stdout)
    assert abs(float(r.group(2)) - np.linalg.norm(xt)) <= 1e-5 * np.linalg.norm(xt)        # errNorm = ||x - 0||
    assert abs(float(r.group(2)) - host["err"]) <= 1e-5 * host["err"]
    row = next(csv.reader(open(hist)))
    assert row[:9] == ["a", "mp", "CGSR", "30", "0", "0", "1e-09", "cuda", "identity"] and int(row[9]) == host["i"] and int(row[10]) == host["iters"]
    assert len(row) == 15 and float(row[13]) >= 0 and float(row[14]) > 0
