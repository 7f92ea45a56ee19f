"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes wrapper over oracle/liboracle.so (built from oracle/oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  It is the checker for the CUDA path, never the thing measured or shipped.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MODES = {"mixed": 0, "baseline": 1, "single-prec": 2, "single": 3}
ORTHS = {"cgs": 0, "mgs": 1, "cgsr": 2}
CONVS = {"base": 0, "relprecres": 1, "repeat": 2, "orthloss": 3}
PRECS = {"identity": 0, "jacobi": 1, "ilu_jacobi": 2}


class Stats(C.Structure):
    _fields_ = [("status", C.c_int64), ("total_iters", C.c_int64), ("total_restarts", C.c_int64),
                ("outer_i", C.c_int64), ("rel_prec_res", C.c_double), ("b_norm", C.c_double),
                ("Minvb_norm", C.c_double), ("A_norm", C.c_double), ("n_hist_inner", C.c_int64),
                ("n_hist_outer", C.c_int64)]


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        L = _LIB
        L.orc_lap2d_nnz.restype = C.c_int64
        L.orc_cd27_nnz.restype = C.c_int64
        L.orc_powerlaw_rowmap.restype = C.c_int64
        L.orc_partition_local.restype = C.c_int64
        L.orc_dot_f32.restype = C.c_float
        L.orc_dot_f64.restype = C.c_double
        L.orc_nrm2_f32.restype = C.c_float
        L.orc_nrm2_f64.restype = C.c_double
        L.orc_givens_step_f32.restype = C.c_double
        L.orc_givens_step_f64.restype = C.c_double
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f(x):
    return C.c_float(x)


def _d(x):
    return C.c_double(x)


def _i64(x):
    return C.c_int64(int(x))


def num_threads():
    return lib().orc_num_threads()


def rand_vect(n, seed=42):
    out = np.empty(n, dtype=np.float64)
    lib().orc_rand_vect(_i64(n), C.c_uint32(seed), _p(out))
    return out


# ---- generators: return (row_map int32[n+1], inds int32[nnz], vals float64[nnz]) -------------------
def gen_lap2d(N):
    n, nnz = N * N, lib().orc_lap2d_nnz(_i64(N))
    rm, ind, val = np.empty(n + 1, np.int32), np.empty(nnz, np.int32), np.empty(nnz, np.float64)
    lib().orc_gen_lap2d(_i64(N), _p(rm), _p(ind), _p(val))
    return rm, ind, val


def gen_cd27(N):
    n, nnz = N ** 3, lib().orc_cd27_nnz(_i64(N))
    rm, ind, val = np.empty(n + 1, np.int32), np.empty(nnz, np.int32), np.empty(nnz, np.float64)
    lib().orc_gen_cd27(_i64(N), _p(rm), _p(ind), _p(val))
    return rm, ind, val


def gen_powerlaw(n, seed=7, lmin=2, gmax=15):
    rm = np.empty(n + 1, np.int32)
    nnz = lib().orc_powerlaw_rowmap(_i64(n), C.c_uint64(seed), C.c_int(lmin), C.c_int(gmax), _p(rm))
    ind, val = np.empty(nnz, np.int32), np.empty(nnz, np.float64)
    lib().orc_gen_powerlaw(_i64(n), C.c_uint64(seed), C.c_int(lmin), C.c_int(gmax), _p(rm), _p(ind), _p(val))
    return rm, ind, val


def gen(spec):
    """spec: 'lap2d:N' | 'cd27:N' | 'powerlaw:n[:seed[:lmin[:gmax]]]'"""
    kind, *args = spec.split(":")
    args = [int(a) for a in args]
    return {"lap2d": gen_lap2d, "cd27": gen_cd27, "powerlaw": gen_powerlaw}[kind](*args)


# ---- ops -------------------------------------------------------------------------------------------
def _sfx(dt):
    return {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64"}[np.dtype(dt)]


def _sc(dt, v):
    return _f(v) if np.dtype(dt) == np.float32 else _d(v)


def spmv(rm, ind, val, alpha, x, beta, y):
    """y = alpha*A*x + beta*y (in place on y); kernels.hpp:159-160"""
    n = len(rm) - 1
    getattr(lib(), "orc_spmv_" + _sfx(val.dtype))(C.c_int(n), _p(rm), _p(ind), _p(val), _sc(val.dtype, alpha),
                                                   _p(x), _sc(val.dtype, beta), _p(y))
    return y


def dot(x, y):
    return getattr(lib(), "orc_dot_" + _sfx(x.dtype))(_i64(len(x)), _p(x), _p(y))


def nrm2(x):
    return getattr(lib(), "orc_nrm2_" + _sfx(x.dtype))(_i64(len(x)), _p(x))


def axpy(alpha, x, y):
    getattr(lib(), "orc_axpy_" + _sfx(x.dtype))(_i64(len(x)), _sc(x.dtype, alpha), _p(x), _p(y))
    return y


def naxpy(alpha, x, y):
    getattr(lib(), "orc_naxpy_" + _sfx(x.dtype))(_i64(len(x)), _sc(x.dtype, alpha), _p(x), _p(y))
    return y


def scal(alpha, x):
    y = np.empty_like(x)
    getattr(lib(), "orc_scal_" + _sfx(x.dtype))(_i64(len(x)), _sc(x.dtype, alpha), _p(x), _p(y))
    return y


def gemv(trans, M, ncols, alpha, x, beta, y):
    """M: Fortran-ordered (nrows, >=ncols) array; uses the first ncols columns."""
    assert M.flags.f_contiguous
    nr, ld = M.shape[0], M.strides[1] // M.itemsize
    getattr(lib(), "orc_gemv_" + _sfx(M.dtype))(C.c_int(1 if trans else 0), _i64(nr), _i64(ncols), _sc(M.dtype, alpha),
                                                 _p(M), _i64(ld), _p(x), _sc(M.dtype, beta), _p(y))
    return y


def trsv_upper(A, n, x):
    assert A.flags.f_contiguous
    ld = A.strides[1] // A.itemsize
    getattr(lib(), "orc_trsv_upper_" + _sfx(A.dtype))(_i64(n), _p(A), _i64(ld), _p(x))
    return x


def rotg(a, b, dtype=np.float32):
    arr = [np.array([v], dtype=dtype) for v in (a, b, 0, 0)]
    getattr(lib(), "orc_rotg_" + _sfx(dtype))(*[_p(v) for v in arr])
    return tuple(v[0] for v in arr)  # (r, 0, c, s)


def givens_step(k, h, cs, sn, s):
    """h: Fortran (m+1, m); applies k old rotations to column k, makes the new one, rotates s. Returns |s[k+1]|."""
    ldh = h.strides[1] // h.itemsize
    return getattr(lib(), "orc_givens_step_" + _sfx(h.dtype))(_i64(k), _p(h), _i64(ldh), _p(cs), _p(sn), _p(s))


def add_vector(orth, V, k, w):
    """GS::add_vector (Orthogonalization.hpp:51-60). V: Fortran (n, >=k+2); w updated in place; returns hcol[k+2]."""
    assert V.flags.f_contiguous and V.shape[1] >= k + 2 and V.strides[1] // V.itemsize == V.shape[0]
    hcol = np.zeros(k + 2, dtype=V.dtype)
    getattr(lib(), "orc_add_vector_" + _sfx(V.dtype))(C.c_int(ORTHS[orth]), _i64(V.shape[0]), _i64(k), _p(V), _p(w), _p(hcol))
    return hcol


def cast(x, dtype):
    y = np.empty(len(x), dtype=dtype)
    name = "orc_cast_f64_f32" if x.dtype == np.float64 else "orc_cast_f32_f64"
    getattr(lib(), name)(_i64(len(x)), _p(x), _p(y))
    return y


# ---- solver ----------------------------------------------------------------------------------------
def gmres(rm, ind, val64, b, x0=None, mode="mixed", orth="cgsr", conv="base", prec="identity", rlen=50, tol=1e-6,
          rtol=0.0, max_restarts=1000000, hist_cap=None, jacobi_steps=1):
    n = len(rm) - 1
    x = np.zeros(n, np.float64) if x0 is None else np.array(x0, np.float64)
    st = Stats()
    cap_outer = min(int(max_restarts) + 2, 100000)
    cap_inner = hist_cap if hist_cap is not None else min(cap_outer * int(rlen), 4000000)
    hi = np.zeros(max(cap_inner, 1), np.float64)
    ho = np.zeros(4 * cap_outer, np.float64)
    lib().orc_gmres2(C.c_int(MODES[mode]), C.c_int(ORTHS[orth]), C.c_int(CONVS[conv]), C.c_int(PRECS[prec]), C.c_int(jacobi_steps), _i64(rlen),
                    _d(tol), _d(rtol), _i64(max_restarts), C.c_int(n), _p(rm), _p(ind), _p(val64), _p(b), _p(x),
                    C.byref(st), _p(hi), _i64(cap_inner), _p(ho), _i64(cap_outer))
    res = {f: getattr(st, f) for f, _ in Stats._fields_}
    res["hist_inner"] = hi[:min(st.n_hist_inner, cap_inner)].copy()
    res["hist_outer"] = ho[:4 * min(st.n_hist_outer, cap_outer)].reshape(-1, 4).copy()
    res["x"] = x
    return res


# ---- ILU(0) + Jacobi sweeps (restated; see oracle.cpp) ----
def ilu0(rm, ind, val64, eps_is_float=True):
    out = np.empty(len(val64), np.float64)
    lib().orc_ilu0(C.c_int(len(rm) - 1), _p(rm), _p(ind), _p(val64), C.c_int(int(eps_is_float)), _p(out))
    return out


def ilu_jacobi_apply(rm, ind, ilu_vals, steps, x):
    """ilusv_jacobi (kernels.hpp:227-248) in x's precision, in place"""
    getattr(lib(), "orc_ilu_jacobi_apply_" + _sfx(x.dtype))(C.c_int(len(rm) - 1), _p(rm), _p(ind), _p(ilu_vals), C.c_int(steps), _p(x))
    return x


def ilu_jacobi_mv(rm, ind, ilu_vals, lower, alpha, x, beta, y):
    getattr(lib(), "orc_ilu_jacobi_mv_" + _sfx(x.dtype))(C.c_int(len(rm) - 1), _p(rm), _p(ind), _p(ilu_vals), C.c_int(int(lower)), _sc(x.dtype, alpha), _p(x),
                                                          _sc(x.dtype, beta), _p(y))
    return y


# ---- partition -------------------------------------------------------------------------------------
def partition_bounds(n, P):
    b = np.empty(P + 1, np.int64)
    lib().orc_partition_bounds(_i64(n), C.c_int(P), _p(b))
    return b


def partition_bounds_nnz(rm, P):
    b = np.empty(P + 1, np.int64)
    lib().orc_partition_bounds_nnz(_i64(len(rm) - 1), C.c_int(P), _p(rm), _p(b))
    return b


def partition_local_range(lo, hi, rm, ind):
    """(halo_cols, local_inds) of the slab of rows [lo, hi) (any split points)"""
    lib().orc_partition_local_range.restype = C.c_int64
    nh = lib().orc_partition_local_range(_i64(lo), _i64(hi), _p(rm), _p(ind), None, None)
    halo = np.empty(nh, np.int64)
    li = np.empty(int(rm[hi] - rm[lo]), np.int32)
    lib().orc_partition_local_range(_i64(lo), _i64(hi), _p(rm), _p(ind), _p(halo), _p(li))
    return halo, li


def partition_local(n, P, r, rm, ind):
    """returns (halo_cols int64[nh] global ids ascending, local_inds int32[nnz_local])"""
    nh = lib().orc_partition_local(_i64(n), C.c_int(P), C.c_int(r), _p(rm), _p(ind), None, None)
    b = partition_bounds(n, P)
    nnz_l = int(rm[b[r + 1]] - rm[b[r]])
    halo = np.empty(nh, np.int64)
    li = np.empty(nnz_l, np.int32)
    lib().orc_partition_local(_i64(n), C.c_int(P), C.c_int(r), _p(rm), _p(ind), _p(halo), _p(li))
    return halo, li


SELL_G, SELL_CHUNK, SELL_SIGMA = 4, 256, 4096


def sell_rows(rm, sigma_mode):
    """lane table of the packed layout of csrc/sell.cu: (start, len, out) per lane.  PLAIN: lane = row.  SIGMA: rows longer
    than SELL_CHUNK nonzeros are cut into pieces of at most SELL_CHUNK ("virtual rows", in (row, piece) order), the virtual rows
    are sorted by length descending (ties: index ascending) inside windows of SELL_SIGMA; out = the row, or -1 - (piece number
    among the pieces of all cut rows) for a piece of a cut row.  Also returns (split_rows, chunk_base)."""
    n = len(rm) - 1
    lens = np.diff(rm).astype(np.int64)
    if not sigma_mode:
        return rm[:-1].astype(np.int32), lens.astype(np.int32), np.arange(n, dtype=np.int32), np.zeros(0, np.int32), np.zeros(1, np.int32)
    nch = np.maximum(1, (lens + SELL_CHUNK - 1) // SELL_CHUNK)
    vrow = np.repeat(np.arange(n), nch)                          # row of every virtual row
    first = np.zeros(n + 1, np.int64); first[1:] = np.cumsum(nch)
    piece = np.arange(len(vrow)) - first[vrow]
    vstart = rm[vrow].astype(np.int64) + piece * SELL_CHUNK
    vlen = np.minimum(SELL_CHUNK, lens[vrow] - piece * SELL_CHUNK)
    split = nch > 1
    split_rows = np.nonzero(split)[0].astype(np.int32)
    cb = np.zeros(len(split_rows) + 1, np.int64); cb[1:] = np.cumsum(nch[split])
    chunk_of_row = np.zeros(n, np.int64); chunk_of_row[split_rows] = cb[:-1]
    vout = np.where(split[vrow], -1 - (chunk_of_row[vrow] + piece), vrow)
    order = np.concatenate([w0 + np.lexsort((np.arange(min(SELL_SIGMA, len(vrow) - w0)), -vlen[w0:w0 + SELL_SIGMA]))
                            for w0 in range(0, len(vrow), SELL_SIGMA)]) if len(vrow) else np.zeros(0, np.int64)
    return vstart[order].astype(np.int32), vlen[order].astype(np.int32), vout[order].astype(np.int32), split_rows, cb.astype(np.int32)


def sell_pack(rm, ind, val, sigma_mode=False):
    """Restatement (numpy, test infrastructure) of the packed sliced-ELL layout of csrc/sell.cu, DESIGN.md §2: 32-lane slices,
    slice length L = longest lane (rounded up to a multiple of 4 when that pads <= 10 %), the first floor(L/4)*4 positions in
    groups of 4 per lane, the L % 4 trailing positions one per lane; short lanes padded with their first column and value 0.
    fp64 values store each group as two half-groups ([lane][2] twice).  Returns (slice_off[nslices + 1] int64, inds int32, vals)."""
    G = SELL_G
    start, lens, _, _, _ = sell_rows(rm, sigma_mode)
    nl = len(start)
    ns = (nl + 31) // 32
    lp = np.zeros(ns * 32, np.int64); lp[:nl] = lens
    sp_ = np.zeros(ns * 32, np.int64); sp_[:nl] = start
    L = lp.reshape(ns, 32).max(axis=1) if ns else np.zeros(0, np.int64)
    rem = L % G
    up = (rem > 0) & ((G - rem) * 10 <= L)
    L = np.where(up, L + G - rem, L)
    off = np.zeros(ns + 1, np.int64)
    off[1:] = np.cumsum(L * 32)
    sind = np.zeros(int(off[-1]), np.int32)
    sval = np.zeros(int(off[-1]), val.dtype)
    f64 = np.dtype(val.dtype) == np.float64
    for s in range(ns):
        Ls = int(L[s]); ng = Ls // G; o = int(off[s])
        p = np.arange(Ls)
        in_grp = p < ng * G
        for lane in range(32):
            ln, rs = int(lp[s * 32 + lane]), int(sp_[s * 32 + lane])
            c = np.full(Ls, ind[rs] if ln > 0 else 0, np.int32)
            v = np.zeros(Ls, val.dtype)
            c[:ln] = ind[rs:rs + ln]
            v[:ln] = val[rs:rs + ln]
            tail = ng * 32 * G + (p - ng * G) * 32 + lane
            sind[o + np.where(in_grp, (p // G) * 32 * G + lane * G + p % G, tail)] = c
            if f64:
                sval[o + np.where(in_grp, (p // G) * 32 * G + ((p % G) // 2) * 64 + lane * 2 + p % 2, tail)] = v
            else:
                sval[o + np.where(in_grp, (p // G) * 32 * G + lane * G + p % G, tail)] = v
    return off, sind, sval
