"""CPU tests of the oracle itself: known answers, generator invariants, BLAS semantics against numpy/scipy in fp64,
and the restated GMRES drivers against scipy and the stopping rule they implement."""
import numpy as np
import pytest
import scipy.linalg
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from util import EPS, problem, summation_bound


def csr(rm, ind, val):
    n = len(rm) - 1
    return sp.csr_matrix((val, ind, rm), shape=(n, n))


def test_rand_vect_known_answer(orc):
    # libstdc++ mt19937(42) + uniform_real_distribution<float>, one 32-bit draw per value
    x = orc.rand_vect(6, 42)
    assert [float(v).hex() for v in x[:3]] == ["0x1.7f87720000000p-2", "0x1.97d47c0000000p-1", "0x1.e6c4060000000p-1"]
    np.testing.assert_allclose(x, [0.37454012, 0.796543002, 0.95071429, 0.183434784, 0.731993914, 0.779690981], rtol=3e-8)
    raw = np.array([1608637542, 3421126067, 4083286876, 787846414, 3143890026, 3348747335], dtype=np.float64)
    np.testing.assert_array_equal(x, (raw.astype(np.float32) / np.float32(2.0 ** 32)).astype(np.float64))


@pytest.mark.parametrize("spec", ["lap2d:1", "lap2d:2", "lap2d:17", "cd27:1", "cd27:2", "cd27:7", "powerlaw:300", "powerlaw:5000:11:2:8"])
def test_generators_canonical_form(orc, spec):
    rm, ind, val = orc.gen(spec)
    n = len(rm) - 1
    assert rm[0] == 0 and rm[-1] == len(ind) == len(val)
    assert np.all(np.diff(rm) >= 1)  # LoadMatrix.hpp:62-66: every row has its diagonal
    for r in range(n):
        cols = ind[rm[r]:rm[r + 1]]
        assert np.all(np.diff(cols) > 0), "columns ascending and distinct (LoadMatrix.hpp:128-145)"
        assert r in cols
        assert cols.min() >= 0 and cols.max() < n
    # fp32-representable values
    np.testing.assert_array_equal(val, val.astype(np.float32).astype(np.float64))
    A = csr(rm, ind, val)
    d = A.diagonal()
    off = np.asarray(abs(A).sum(axis=1)).ravel() - np.abs(d)
    kind = spec.split(":")[0]
    if kind == "lap2d":
        N = int(spec.split(":")[1])
        assert len(val) == 5 * N * N - 4 * N
        assert (A != A.T).nnz == 0 and np.all(d == 4)
    elif kind == "cd27":
        N = int(spec.split(":")[1])
        assert len(val) == (3 * N - 2) ** 3
        assert np.all(d == 26) and np.all(off <= 26)
        if N > 1:
            assert (A != A.T).nnz > 0  # convection makes it nonsymmetric
    else:
        np.testing.assert_array_equal(d, 1.0 + off)  # strict dominance by exactly 1, exact in fp64 and fp32
        assert (A != A.T).nnz > 0


def test_powerlaw_row_lengths_are_skewed(orc):
    rm, _, _ = orc.gen("powerlaw:200000")
    lens = np.diff(rm)
    assert lens.min() >= 3 and lens.max() > 50 * np.median(lens)
    assert 15 < lens.mean() < 40  # ~25 nnz/row like config 4 (8 M rows, ~200 M nnz)


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_blas_ops_against_numpy(orc, dt):
    rng = np.random.default_rng(0)
    n, k = 5003, 7
    x = rng.standard_normal(n).astype(dt)
    y = rng.standard_normal(n).astype(dt)
    x64, y64 = x.astype(np.float64), y.astype(np.float64)
    assert abs(orc.dot(x, y) - x64 @ y64) <= summation_bound(np.abs(x64 * y64).sum(), n, dt)
    assert abs(orc.nrm2(x) - np.linalg.norm(x64)) <= 4 * EPS[np.dtype(dt)] * np.linalg.norm(x64) * np.sqrt(n)
    a = dt(0.37)
    # axpy / naxpy are single fused multiply-adds: compare with the exactly rounded result via float128-free identity
    got = orc.axpy(a, x, y.copy())
    ref = (np.float64(a) * x64 + y64)
    assert np.max(np.abs(got - ref)) <= 2 * EPS[np.dtype(dt)] * np.max(np.abs(ref) + 1)
    got = orc.naxpy(a, x, y.copy())
    ref = (y64 - np.float64(a) * x64)
    assert np.max(np.abs(got - ref)) <= 2 * EPS[np.dtype(dt)] * np.max(np.abs(ref) + 1)
    np.testing.assert_array_equal(orc.scal(a, x), a * x)
    # gemv, column-major with ld = nrows
    M = np.asfortranarray(rng.standard_normal((n, k)).astype(dt))
    h = np.zeros(k, dt)
    orc.gemv(True, M, k, 1.0, x, 0.0, h)
    ref = M.astype(np.float64).T @ x64
    bound = summation_bound((np.abs(M.astype(np.float64)) * np.abs(x64)[:, None]).sum(axis=0).max(), n, dt)
    assert np.max(np.abs(h - ref)) <= bound
    yy = y.copy()
    orc.gemv(False, M, k, -1.0, h, 1.0, yy)
    ref = y64 - M.astype(np.float64) @ h.astype(np.float64)
    assert np.max(np.abs(yy - ref)) <= 16 * EPS[np.dtype(dt)] * (np.abs(M.astype(np.float64)) @ np.abs(h.astype(np.float64)) + np.abs(y64)).max()
    # casts are round-to-nearest
    if dt == np.float64:
        np.testing.assert_array_equal(orc.cast(x, np.float32), x.astype(np.float32))
    else:
        np.testing.assert_array_equal(orc.cast(x, np.float64), x.astype(np.float64))


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_givens_and_trsv(orc, dt):
    # rotg: BLAS sign convention, then b := 0 (kernels_mkl.cpp:207-219)
    for a, b in [(3.0, 4.0), (-3.0, 4.0), (4.0, -3.0), (-4.0, -3.0), (0.0, 2.0), (2.0, 0.0), (0.0, 0.0)]:
        r, z, c, s = orc.rotg(a, b, dt)
        assert z == 0
        if a == 0 and b == 0:
            assert (r, c, s) == (0, 1, 0)
            continue
        roe = a if abs(a) > abs(b) else b
        assert np.sign(r) == np.sign(roe)
        assert abs(abs(r) - np.hypot(a, b)) <= 4 * EPS[np.dtype(dt)] * np.hypot(a, b)
        assert abs(c * r - a) <= 4 * EPS[np.dtype(dt)] * abs(r) and abs(s * r - b) <= 4 * EPS[np.dtype(dt)] * abs(r)
    # a full Givens QR of a Hessenberg matrix reproduces |R| from numpy's QR and the least-squares residual
    rng = np.random.default_rng(1)
    m = 12
    H = np.asfortranarray(np.triu(rng.standard_normal((m + 1, m)), -1).astype(dt))
    H0 = H.astype(np.float64).copy()
    cs, sn, s = np.zeros(m + 1, dt), np.zeros(m + 1, dt), np.zeros(m + 1, dt)
    s[0] = 1.5
    res = [orc.givens_step(k, H, cs, sn, s) for k in range(m)]
    R = np.linalg.qr(H0, mode="r")
    tol = 200 * EPS[np.dtype(dt)] * np.abs(H0).max() * m
    assert np.max(np.abs(np.abs(np.triu(H[:m, :m])) - np.abs(R))) <= tol
    assert np.all(np.tril(H[:m + 1, :m], -1) == 0)  # rotg zeroes the subdiagonal (b := 0)
    e1 = np.zeros(m + 1); e1[0] = 1.5
    lsq = np.linalg.lstsq(H0, e1, rcond=None)
    assert abs(res[-1] - np.linalg.norm(H0 @ lsq[0] - e1)) <= tol
    # trsv Upper/NoTrans/NonUnit in place on s (gmres.cpp:285-288)
    y = s[:m].copy()
    orc.trsv_upper(H, m, y)
    ref = scipy.linalg.solve_triangular(np.triu(H[:m, :m]).astype(np.float64), s[:m].astype(np.float64))
    assert np.max(np.abs(y - ref)) <= 1e4 * EPS[np.dtype(dt)] * np.abs(ref).max() * np.linalg.cond(np.triu(H[:m, :m]).astype(np.float64))
    np.testing.assert_allclose(y, lsq[0], rtol=0, atol=1e5 * EPS[np.dtype(dt)] * np.abs(lsq[0]).max() * np.linalg.cond(H0))


@pytest.mark.parametrize("spec", ["lap2d:33", "cd27:9", "powerlaw:3000"])
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_spmv_against_scipy(orc, spec, dt):
    rm, ind, val = orc.gen(spec)
    n = len(rm) - 1
    A = csr(rm, ind, val)
    rng = np.random.default_rng(3)
    x = rng.standard_normal(n).astype(dt)
    y = rng.standard_normal(n).astype(dt)
    v = val.astype(dt)
    got = orc.spmv(rm, ind, v, -1.0, x, 1.0, y.copy())
    ref = y.astype(np.float64) - A @ x.astype(np.float64)
    bound = summation_bound((abs(A) @ np.abs(x.astype(np.float64)) + np.abs(y)).max(), int(np.diff(rm).max()), dt)
    assert np.max(np.abs(got - ref)) <= bound
    got0 = orc.spmv(rm, ind, v, 1.0, x, 0.0, np.full(n, np.nan, dt))  # beta = 0 ignores y (even NaN)
    assert np.all(np.isfinite(got0))


@pytest.mark.parametrize("orth", ["cgs", "mgs", "cgsr"])
def test_add_vector_orthogonalises(orc, orth):
    rng = np.random.default_rng(5)
    n, k = 4001, 9
    Q, _ = np.linalg.qr(rng.standard_normal((n, k + 1)))
    V = np.zeros((n, k + 2), np.float32, order="F")
    V[:, :k + 1] = Q.astype(np.float32)
    w0 = rng.standard_normal(n).astype(np.float32)
    w = w0.copy()
    h = orc.add_vector(orth, V, k, w)
    V64 = V.astype(np.float64)
    # Arnoldi relation: w0 = V[:, :k+2] h ; new column is unit and orthogonal to the old ones
    assert np.linalg.norm(V64 @ h.astype(np.float64) - w0) <= 2e-6 * np.linalg.norm(w0)
    assert abs(np.linalg.norm(V64[:, k + 1]) - 1) <= 1e-6
    lim = 2e-6 if orth != "cgs" else 2e-5
    assert np.max(np.abs(V64[:, :k + 1].T @ V64[:, k + 1])) <= lim


@pytest.mark.parametrize("spec,rlen", [("lap2d:24", 30), ("cd27:8", 20), ("powerlaw:1500", 20)])
@pytest.mark.parametrize("mode", ["mixed", "baseline", "single-prec", "single"])
def test_gmres_reaches_reference_criterion_and_matches_scipy(orc, spec, rlen, mode):
    rm, ind, val, xt, b = problem(orc, spec)
    A = csr(rm, ind, val)
    tol = 1e-9 if mode in ("mixed", "baseline") else 1e-6
    r = orc.gmres(rm, ind, val, b, mode=mode, orth="cgsr", rlen=rlen, tol=tol, max_restarts=400)
    assert r["status"] == 1
    assert r["total_iters"] == (r["total_restarts"] - 1) * rlen  # base Convergence restarts only at k = rlen
    x = r["x"]
    # the stopping rule of IterUtil.hpp:42-51 holds for the returned x (recomputed independently in fp64)
    res = np.linalg.norm(b - A @ x)
    A_norm = np.linalg.norm(val.astype(np.float32).astype(np.float64))
    crit = res / (np.linalg.norm(b) + A_norm * np.linalg.norm(x))
    slack = 1.0 if mode in ("mixed", "baseline") else 30.0  # fp32 solvers see an fp32-rounded residual
    assert crit <= tol * slack * 1.001
    # same answer as an independent GMRES
    xs, info = spla.gmres(A, b, rtol=1e-12, restart=rlen, maxiter=2000)
    assert info == 0
    assert np.linalg.norm(x - xs) <= 1e3 * max(tol, 1e-7) * (A_norm / np.abs(A.diagonal()).min()) * np.linalg.norm(xs)
    # history bookkeeping
    assert len(r["hist_inner"]) == r["total_iters"] and len(r["hist_outer"]) == r["total_restarts"]
    assert np.all(np.diff(r["hist_inner"][:rlen]) <= 1e-6 * r["hist_inner"][0])  # Arnoldi residual is non-increasing in a cycle


def test_gmres_restart_policies(orc):
    rm, ind, val, xt, b = problem(orc, "lap2d:24")
    base = orc.gmres(rm, ind, val, b, mode="mixed", rlen=40, tol=1e-9, max_restarts=400)
    rel = orc.gmres(rm, ind, val, b, mode="mixed", rlen=40, tol=1e-9, conv="relprecres", rtol=1e-2, max_restarts=4000)
    rep = orc.gmres(rm, ind, val, b, mode="mixed", rlen=40, tol=1e-9, conv="repeat", rtol=1e-2, max_restarts=4000)
    lo = orc.gmres(rm, ind, val, b, mode="mixed", rlen=40, tol=1e-9, conv="orthloss", rtol=1e-3, max_restarts=4000)
    for r in (base, rel, rep, lo):
        assert r["status"] == 1
    # RelPrecRes restarts as soon as the Arnoldi residual improved by rtol: cycles are shorter than rlen
    assert rel["total_restarts"] > base["total_restarts"]
    # RepeatIteration: all cycles after the first have the first cycle's length
    first = np.argmax(rep["hist_inner"] / rep["hist_outer"][0, 2] * rep["Minvb_norm"] <= 1e-2) + 1
    assert (rep["total_iters"] - first) % first == 0
    # max_restarts exceeded => aborted (IterUtil.hpp:44-45)
    ab = orc.gmres(rm, ind, val, b, mode="mixed", rlen=5, tol=1e-14, max_restarts=3)
    assert ab["status"] == 3 and ab["total_restarts"] == 4


def test_jacobi_preconditioner(orc):
    rm, ind, val, xt, b = problem(orc, "powerlaw:1500")
    r0 = orc.gmres(rm, ind, val, b, mode="mixed", rlen=20, tol=1e-9, max_restarts=400)
    r1 = orc.gmres(rm, ind, val, b, mode="mixed", rlen=20, tol=1e-9, prec="jacobi", max_restarts=400)
    assert r0["status"] == 1 and r1["status"] == 1
    assert r1["total_iters"] <= r0["total_iters"]
    np.testing.assert_allclose(r1["x"], xt, rtol=0, atol=1e-4)


@pytest.mark.parametrize("spec", ["lap2d:9", "cd27:6", "powerlaw:400"])
@pytest.mark.parametrize("P", [1, 2, 3, 8])
def test_partition_index_sets(orc, spec, P):
    rm, ind, val = orc.gen(spec)
    n = len(rm) - 1
    b = orc.partition_bounds(n, P)
    assert b[0] == 0 and b[-1] == n and np.all(np.diff(b) >= 0)
    np.testing.assert_array_equal(b, [(r * n) // P for r in range(P + 1)])
    x = np.random.default_rng(0).standard_normal(n)
    y = csr(rm, ind, val) @ x
    for r in range(P):
        halo, li = orc.partition_local(n, P, r, rm, ind)
        lo, hi = b[r], b[r + 1]
        assert np.all(np.diff(halo) > 0) and np.all((halo < lo) | (halo >= hi))
        # the local slab applied to [x_local ; x_halo] reproduces the global rows
        xl = np.concatenate([x[lo:hi], x[halo]])
        rml = rm[lo:hi + 1] - rm[lo]
        yl = sp.csr_matrix((val[rm[lo]:rm[hi]], li, rml), shape=(hi - lo, len(xl))) @ xl
        np.testing.assert_allclose(yl, y[lo:hi], rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("spec,sigma,dt", [("cd27:5", False, np.float32), ("lap2d:9", False, np.float32), ("lap2d:9", False, np.float64),
                                           ("powerlaw:300", False, np.float64), ("powerlaw:9000", True, np.float32), ("powerlaw:9000", True, np.float64)])
def test_packed_layout_restatement_is_a_permutation_of_the_csr_entries(orc, spec, sigma, dt):
    """oracle.sell_pack / sell_rows (the layout the GPU test compares the library's packed arrays with): every CSR entry appears
    exactly once at the documented position, padding carries value 0 and a valid column, y = A x through the packed arrays
    (lane sums scattered through the lane table, pieces of cut rows added) is exact"""
    G = 4
    rm, ind, val = orc.gen(spec)
    val = val.astype(dt)
    n = len(rm) - 1
    off, si, sv = orc.sell_pack(rm, ind, val, sigma_mode=sigma)
    lstart, llen, lout, split_rows, chunk_base = orc.sell_rows(rm, sigma)
    nl = len(lstart)
    assert off[0] == 0 and np.all(np.diff(off) % 32 == 0) and len(si) == off[-1] == len(sv)
    assert np.count_nonzero(sv) == np.count_nonzero(val) and si.min() >= 0 and si.max() < n
    assert llen.sum() == len(ind) and (not sigma or llen.max() <= 256)
    f64 = dt == np.float64
    x = np.arange(1, n + 1, dtype=np.float64)
    y = np.zeros(n)
    partial = np.zeros(int(chunk_base[-1]))
    for s in range(len(off) - 1):
        L = int(off[s + 1] - off[s]) // 32
        ng = L // G
        for lane in range(32):
            pos_lane = s * 32 + lane
            if pos_lane >= nl:
                continue
            acc = 0.0
            for p in range(L):
                tail = ng * 32 * G + (p - ng * G) * 32 + lane
                pi = off[s] + ((p // G) * 32 * G + lane * G + p % G if p < ng * G else tail)
                pv = off[s] + (((p // G) * 32 * G + ((p % G) // 2) * 64 + lane * 2 + p % 2 if f64 else (p // G) * 32 * G + lane * G + p % G) if p < ng * G else tail)
                acc += float(sv[pv]) * x[si[pi]]
            o = int(lout[pos_lane])
            if o >= 0:
                y[o] = acc
            else:
                partial[-1 - o] = acc
    for j, r in enumerate(split_rows):
        y[r] = partial[chunk_base[j]:chunk_base[j + 1]].sum()
    A = sp.csr_matrix((val.astype(np.float64), ind, rm), shape=(n, n))
    np.testing.assert_allclose(y, A @ x, rtol=1e-13, atol=1e-9)
    if sigma:
        # sorted by length (descending) inside every window of 4096 lanes; pieces of a cut row keep their order
        for w0 in range(0, nl, 4096):
            assert np.all(np.diff(llen[w0:w0 + 4096]) <= 0)
        assert len(split_rows) == (np.diff(rm) > 256).sum() and off[-1] <= 1.25 * len(ind) + 4096
    # slice lengths: longest lane, rounded up to a multiple of G only when that pads <= 10 %
    lens = np.zeros((len(off) - 1) * 32, np.int64); lens[:nl] = llen
    Lmax = lens.reshape(-1, 32).max(axis=1)
    L = np.diff(off) // 32
    assert np.all(L >= Lmax) and np.all(L - Lmax < G) and np.all((L == Lmax) | ((L - Lmax) * 10 <= Lmax))


@pytest.mark.parametrize("spec", ["lap2d:12", "cd27:6", "powerlaw:800"])
def test_ilu0_restatement_properties(orc, spec):
    """oracle.ilu0 (sequential IKJ, kernels_mkl.cpp:451-484 with diag_inds filled in): L U reproduces A exactly on the pattern of A
    (the defining property of ILU(0)), L is unit lower, pivots keep their sign; IluJacobi with many sweeps converges to the exact
    triangular solves and its first sweep is the first-order Neumann term"""
    rm, ind, val = orc.gen(spec)
    n = len(rm) - 1
    f = orc.ilu0(rm, ind, val, True)
    A, F = csr(rm, ind, val), csr(rm, ind, f)
    L, U = (sp.tril(F, -1) + sp.eye(n)).tocsr(), sp.triu(F).tocsr()
    R = (L @ U - A).tocsr()
    P = A.copy(); P.data[:] = 1
    assert abs(R.multiply(P)).max() <= 64 * EPS[np.dtype(np.float64)] * abs(A).max() * np.diff(rm).max()
    assert np.all(U.diagonal() > 0)
    x0 = np.random.default_rng(1).standard_normal(n)
    exact = spla.spsolve_triangular(U, spla.spsolve_triangular(L, x0, lower=True), lower=False)
    if spec != "powerlaw:800":
        x = orc.ilu_jacobi_apply(rm, ind, f, 80, x0.copy())
        np.testing.assert_allclose(x, exact, rtol=0, atol=1e-9 * np.abs(exact).max())
    # one sweep per triangle = (I - D^-1 S) D^-1 (I - N) x0 with L = I + N, U = D + S
    N, D, S = sp.tril(F, -1), F.diagonal(), sp.triu(F, 1)
    t = x0 - N @ x0
    first = t / D - (S @ (t / D)) / D
    # (kernels.hpp:236-248 with steps = 1: x <- x + (b - L x), then x <- x + D^-1 (b - U x) from x = b: = b + D^-1 b - D^-1 U b)
    t2 = t + (t - U @ t) / D
    x1 = orc.ilu_jacobi_apply(rm, ind, f, 1, x0.copy())
    np.testing.assert_allclose(x1, t2, rtol=0, atol=1e-12 * max(1.0, np.abs(t2).max()))
    del first
    # mv forms
    y0 = np.random.default_rng(2).standard_normal(n)
    np.testing.assert_allclose(orc.ilu_jacobi_mv(rm, ind, f, True, 0.5, x0, -2.0, y0.copy()), -2.0 * y0 + 0.5 * (L @ x0), rtol=0, atol=1e-11 * np.abs(y0).max() * n ** 0.5)
    np.testing.assert_allclose(orc.ilu_jacobi_mv(rm, ind, f, False, 3.0, x0, 7.0, y0.copy()), y0 - U @ x0, rtol=0, atol=1e-11 * (abs(U) @ np.abs(x0) + np.abs(y0)).max())


def test_oracle_gmres_with_ilu_jacobi_converges(orc):
    rm, ind, val, xt, b = problem(orc, "cd27:8")
    r0 = orc.gmres(rm, ind, val, b, mode="mixed", rlen=30, tol=1e-9, max_restarts=50)
    r2 = orc.gmres(rm, ind, val, b, mode="mixed", rlen=30, tol=1e-9, prec="ilu_jacobi", jacobi_steps=2, max_restarts=50)
    r4 = orc.gmres(rm, ind, val, b, mode="baseline", rlen=30, tol=1e-10, prec="ilu_jacobi", jacobi_steps=4, max_restarts=50)
    assert r0["status"] == r2["status"] == r4["status"] == 1
    assert np.linalg.norm(r2["x"] - xt) <= 1e-6 * np.linalg.norm(xt) and np.linalg.norm(r4["x"] - xt) <= 1e-6 * np.linalg.norm(xt)
    # more sweeps = better triangular solves = a faster-falling preconditioned Arnoldi residual (two sweeps alone are a weak
    # preconditioner on this stencil: the Neumann series of L^-1 and U^-1 converges slowly when the off-diagonal row sums are ~0.5)
    assert r4["hist_inner"][10] < r2["hist_inner"][10]
