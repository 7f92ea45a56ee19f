"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes wrapper over oracle/_ref/libref.so: the REFERENCE'S OWN gmres.cpp /
Orthogonalization.hpp / IterUtil.hpp / kernels_mkl.cpp compiled unmodified (oracle/ref.mk) and linked to the oneMKL
inside libtorch_cpu.so.  Used to pin the oracle, to generate tests/golden/*.json, and as bench.py's CPU reference arm.
The .so is built in the container that has /root/reference and travels to the GPU box with the snapshot."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libref.so")
_LIB = None

MODES = {"mixed": 0, "baseline": 1, "single-prec": 2, "single": 3}
ORTHS = {"cgs": 0, "mgs": 1, "cgsr": 2}
CONVS = {"base": 0, "relprecres": 1, "repeat": 2, "orthloss": 3}
PRECS = {"identity": 0, "jacobi": 1}


class RefStats(C.Structure):
    _fields_ = [("status", C.c_int64), ("total_iters", C.c_int64), ("total_restarts", C.c_int64), ("outer_i", C.c_int64),
                ("rel_prec_res", C.c_double), ("res_norm", C.c_double), ("err_norm", C.c_double), ("gmres_seconds", C.c_double),
                ("prec_seconds", C.c_double), ("n_hist_inner", C.c_int64), ("n_hist_outer", C.c_int64)]


def available():
    return os.path.exists(_PATH)


def lib():
    global _LIB
    if _LIB is None:
        import torch  # noqa: F401  (libref.so links libtorch_cpu.so for its embedded MKL)
        _LIB = C.CDLL(_PATH)
    return _LIB


def num_threads():
    return lib().ref_num_threads()


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def gmres(rm, ind, val64, b, x0=None, true_x=None, mode="mixed", orth="cgsr", conv="base", prec="identity", rlen=50, tol=1e-6, rtol=0.0,
          max_restarts=1000000, hist_cap=None):
    n = len(rm) - 1
    rm = np.ascontiguousarray(rm, np.int32); ind = np.ascontiguousarray(ind, np.int32)
    val64 = np.ascontiguousarray(val64, np.float64); b = np.ascontiguousarray(b, np.float64)
    x = np.zeros(n, np.float64) if x0 is None else np.array(x0, np.float64)
    st = RefStats()
    cap_outer = min(int(max_restarts) + 2, 100000)
    cap_inner = hist_cap if hist_cap is not None else min(cap_outer * int(rlen), 4000000)
    hi = np.zeros(max(cap_inner, 1), np.float64)
    ho = np.zeros(4 * cap_outer, np.float64)
    tx = None if true_x is None else np.ascontiguousarray(true_x, np.float64)
    lib().ref_gmres(C.c_int(MODES[mode]), C.c_int(ORTHS[orth]), C.c_int(CONVS[conv]), C.c_int(PRECS[prec]), C.c_int64(rlen), C.c_double(tol),
                    C.c_double(rtol), C.c_int64(max_restarts), C.c_int(n), _p(rm), _p(ind), _p(val64), _p(b), _p(x), _p(tx), C.byref(st), _p(hi),
                    C.c_int64(cap_inner), _p(ho), C.c_int64(cap_outer))
    res = {f: getattr(st, f) for f, _ in RefStats._fields_}
    res["hist_inner"] = hi[:min(st.n_hist_inner, cap_inner)].copy()
    res["hist_outer"] = ho[:4 * min(st.n_hist_outer, cap_outer)].reshape(-1, 4).copy()
    res["x"] = x
    return res


def load_matrix(path):
    """the reference's own LoadMatrix<double>() -> (row_map, inds, vals); raises ValueError with its exception text"""
    n, nnz = C.c_int(), C.c_long()
    err = C.create_string_buffer(256)
    if lib().ref_load_matrix(str(path).encode(), C.byref(n), C.byref(nnz), None, None, None, err, 256) != 0:
        raise ValueError(err.value.decode())
    rm, ind, val = np.empty(n.value + 1, np.int32), np.empty(nnz.value, np.int32), np.empty(nnz.value, np.float64)
    lib().ref_load_matrix(str(path).encode(), C.byref(n), C.byref(nnz), _p(rm), _p(ind), _p(val), err, 256)
    return rm, ind, val


def load_vector(path, col=0):
    """the reference's own LoadVector<double>(file, col) -> float64 array; raises ValueError with its exception text"""
    n = C.c_long()
    err = C.create_string_buffer(256)
    if lib().ref_load_vector(str(path).encode(), C.c_int(col), C.byref(n), None, err, 256) != 0:
        raise ValueError(err.value.decode())
    out = np.empty(n.value, np.float64)
    lib().ref_load_vector(str(path).encode(), C.c_int(col), C.byref(n), _p(out), err, 256)
    return out


def sweep_spmv(rm, ind, val64, is_float, trials=3):
    """seconds per call of the reference's spmv<T,MKL> (one warm pass first)"""
    rm = np.ascontiguousarray(rm, np.int32); ind = np.ascontiguousarray(ind, np.int32); val64 = np.ascontiguousarray(val64, np.float64)
    out = np.zeros(trials, np.float64)
    lib().ref_sweep_spmv(C.c_int(len(rm) - 1), _p(rm), _p(ind), _p(val64), C.c_int(int(is_float)), C.c_int(trials), _p(out))
    return out


def sweep_add_vector(n, m, orth, is_float, ks, trials=3):
    """seconds per call of the reference's GS<T, Kernel<T,MKL>, MKL>::add_vector at every k in ks: array (len(ks), trials)"""
    ks = np.ascontiguousarray(ks, np.int32)
    out = np.zeros(len(ks) * trials, np.float64)
    lib().ref_sweep_add_vector(C.c_int64(n), C.c_int(m), C.c_int(ORTHS[orth]), C.c_int(int(is_float)), _p(ks), C.c_int(len(ks)), C.c_int(trials), _p(out))
    return out.reshape(len(ks), trials)
