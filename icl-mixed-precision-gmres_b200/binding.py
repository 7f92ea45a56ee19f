"""ctypes binding of include/mpgmres_b200.h.  Device operands are torch CUDA tensors (PyTorch is used for device
memory, streams and torch.distributed only); host operands are numpy arrays or pinned torch CPU tensors."""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_LIB = None

MODES = {"mixed": 0, "baseline": 1, "single-prec": 2, "single": 3}
ORTHS = {"cgs": 0, "mgs": 1, "cgsr": 2}
CONVS = {"base": 0, "relprecres": 1, "repeat": 2, "orthloss": 3}
PRECS = {"identity": 0, "jacobi": 1, "ilu_jacobi": 2}


class MpgError(RuntimeError):
    pass


class GmresParams(C.Structure):
    _fields_ = [("mode", C.c_int32), ("orth", C.c_int32), ("conv", C.c_int32), ("prec", C.c_int32),
                ("restart_length", C.c_int64), ("tol", C.c_double), ("restart_tol", C.c_double),
                ("max_restarts", C.c_int64), ("jacobi_steps", C.c_int64)]


class GmresStats(C.Structure):
    _fields_ = [("status", C.c_int64), ("total_iters", C.c_int64), ("total_restarts", C.c_int64),
                ("outer_i", C.c_int64), ("rel_prec_res", C.c_double), ("b_norm", C.c_double),
                ("Minvb_norm", C.c_double), ("A_norm", C.c_double), ("n_hist_inner", C.c_int64),
                ("n_hist_outer", C.c_int64), ("solve_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
                ("launches", C.c_int64), ("h2d_all_ms", C.c_double), ("h2d_bytes", C.c_int64), ("host_overlap", C.c_int64)]


def library_path():
    return os.path.join(_HERE, "lib", "libmpgmres_b200.so")


def header_symbols():
    """every function name declared in include/mpgmres_b200.h"""
    txt = open(os.path.join(_ROOT, "include", "mpgmres_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mpg_[a-z0-9_]+)\s*\(", txt)))


def exported_symbols():
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", library_path()], capture_output=True, text=True, check=True).stdout
    return sorted(set(l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("mpg_")))


def load_library():
    """Load the CUDA backend.  Fails loudly if it has not been built: there is no fallback path."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise MpgError(f"{path} is missing: build it with `python icl-mixed-precision-gmres_b200/build.py` "
                       "(the CUDA extension is the only implementation; there is no CPU fallback)")
    try:  # torch brings libcudart.so.12 into the process; harmless if it is already resolvable
        import torch  # noqa: F401
    except Exception:
        pass
    L = C.CDLL(path)
    L.mpg_version.restype = C.c_char_p
    L.mpg_last_error.restype = C.c_char_p
    L.mpg_last_error.argtypes = [C.c_void_p]
    L.mpg_ctx_stream.restype = C.c_void_p
    L.mpg_launch_count.restype = C.c_int64
    L.mpg_lap2d_nnz.restype = C.c_int64
    L.mpg_cd27_nnz.restype = C.c_int64
    L.mpg_lap2d_nnz.argtypes = [C.c_int64]
    L.mpg_cd27_nnz.argtypes = [C.c_int64]
    L.mpg_host_free.argtypes = [C.c_void_p]
    L.mpg_host_free.restype = None
    _LIB = L
    return L


def read_matrix_market(path):
    """MatrixMarket file -> (row_map int32[n+1], inds int32[nnz], vals float64[nnz]) numpy arrays in the reference's
    LoadMatrix-canonical CSR form (mpg_mm_read_host).  Raises MpgError with the reference's exception text."""
    import numpy as np
    L = load_library()
    n, m, nnz = C.c_int(), C.c_int(), C.c_int64()
    prm, pin, pv = C.POINTER(C.c_int)(), C.POINTER(C.c_int)(), C.POINTER(C.c_double)()
    err = C.create_string_buffer(256)
    rc = L.mpg_mm_read_host(str(path).encode(), C.byref(n), C.byref(m), C.byref(nnz), C.byref(prm), C.byref(pin), C.byref(pv), err, 256)
    if rc != 0:
        raise MpgError(err.value.decode() or f"mpg_mm_read_host failed ({rc})")
    try:
        rm = np.ctypeslib.as_array(prm, shape=(m.value + 1,)).copy()
        ind = np.ctypeslib.as_array(pin, shape=(max(nnz.value, 1),))[:nnz.value].copy()
        val = np.ctypeslib.as_array(pv, shape=(max(nnz.value, 1),))[:nnz.value].copy()
    finally:
        L.mpg_host_free(prm); L.mpg_host_free(pin); L.mpg_host_free(pv)
    return rm, ind, val


def read_matrix_market_slab(path, lo, hi, want_global_rowmap=False):
    """rows [lo, hi) of the canonical CSR of a MatrixMarket file (mpg_mm_read_slab_host: streamed, only the slab is ever held) ->
    (n, nnz_global, row_map_local int32[hi-lo+1], inds int32 (GLOBAL columns), vals float64[, row_map_global int32[n+1]]).
    hi < 0: to the last row.  Equal to rows [lo, hi) of read_matrix_market bit for bit."""
    import numpy as np
    L = load_library()
    n, m, nnzg, nnzl = C.c_int(), C.c_int(), C.c_int64(), C.c_int64()
    prm, pin, pv, pg = C.POINTER(C.c_int)(), C.POINTER(C.c_int)(), C.POINTER(C.c_double)(), C.POINTER(C.c_int)()
    err = C.create_string_buffer(256)
    rc = L.mpg_mm_read_slab_host(str(path).encode(), C.c_int64(lo), C.c_int64(hi), C.byref(n), C.byref(m), C.byref(nnzg), C.byref(nnzl), C.byref(prm),
                                 C.byref(pin), C.byref(pv), C.byref(pg) if want_global_rowmap else None, err, 256)
    if rc != 0:
        raise MpgError(err.value.decode() or f"mpg_mm_read_slab_host failed ({rc})")
    try:
        hi_eff = m.value if hi < 0 else hi
        rm = np.ctypeslib.as_array(prm, shape=(hi_eff - lo + 1,)).copy()
        ind = np.ctypeslib.as_array(pin, shape=(max(nnzl.value, 1),))[:nnzl.value].copy()
        val = np.ctypeslib.as_array(pv, shape=(max(nnzl.value, 1),))[:nnzl.value].copy()
        out = (m.value, nnzg.value, rm, ind, val)
        if want_global_rowmap:
            out = out + (np.ctypeslib.as_array(pg, shape=(m.value + 1,)).copy(),)
    finally:
        L.mpg_host_free(prm); L.mpg_host_free(pin); L.mpg_host_free(pv)
        if want_global_rowmap:
            L.mpg_host_free(pg)
    return out


def read_matrix_market_vector(path, col=0):
    """column `col` of a MatrixMarket array / coordinate file as a float64 numpy array (mpg_mm_read_vector_host = the reference's
    LoadVector, LoadMatrix.hpp:156-233).  Raises MpgError with the reference's exception text."""
    import numpy as np
    L = load_library()
    n, pv = C.c_int64(), C.POINTER(C.c_double)()
    err = C.create_string_buffer(256)
    rc = L.mpg_mm_read_vector_host(str(path).encode(), C.c_int(col), C.byref(n), C.byref(pv), err, 256)
    if rc != 0:
        raise MpgError(err.value.decode() or f"mpg_mm_read_vector_host failed ({rc})")
    try:
        out = np.ctypeslib.as_array(pv, shape=(max(n.value, 1),))[:n.value].copy()
    finally:
        L.mpg_host_free(pv)
    return out


def _ptr(t):
    """device/host pointer of a torch tensor, numpy array, int or None"""
    if t is None:
        return C.c_void_p(0)
    if isinstance(t, int):
        return C.c_void_p(t)
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


def _sfx(t):
    import torch
    return {torch.float32: "f32", torch.float64: "f64"}[t.dtype]


def _sc(t, v):
    import torch
    return C.c_float(v) if t.dtype == torch.float32 else C.c_double(v)


class CSR:
    """mpg_csr handle: CSR structure + SpMV plan (values are passed per call, fp32 and fp64 share the structure)."""

    def __init__(self, ctx, row_map, inds, ncols=None):
        self.ctx, self.row_map, self.inds = ctx, row_map, inds
        self.nrows = row_map.numel() - 1
        self.ncols = self.nrows if ncols is None else ncols
        self.nnz = inds.numel()
        self.h = C.c_void_p()
        ctx._chk(ctx.L.mpg_csr_create(ctx.h, C.c_int(self.nrows), C.c_int(self.ncols), C.c_int64(self.nnz), _ptr(row_map),
                                      _ptr(inds), C.byref(self.h)))

    def __del__(self):
        try:
            if self.h:
                self.ctx.L.mpg_csr_destroy(self.h)
                self.h = None
        except Exception:
            pass


class Packed:
    """mpg_packed handle: sliced-ELL copy of a CSR matrix + values (None-like when the structure does not pack well: .h is null)"""

    def __init__(self, ctx, A, vals):
        self.ctx, self.A, self.sfx = ctx, A, _sfx(vals)
        self.h = C.c_void_p()
        ctx._chk(getattr(ctx.L, "mpg_pack_create_" + self.sfx)(ctx.h, A.h, _ptr(vals), C.byref(self.h)))

    def __bool__(self):
        return bool(self.h)

    def update(self, vals):
        self.ctx._chk(getattr(self.ctx.L, "mpg_pack_update_" + self.sfx)(self.ctx.h, self.h, _ptr(vals)))

    def arrays(self):
        """(G, slice_off, inds, vals) copied to the host as numpy arrays (tests: bit-exact layout check against the oracle)"""
        import numpy as np
        G, ns, tot = C.c_int(), C.c_int(), C.c_int64()
        po, pi, pv = C.c_void_p(), C.c_void_p(), C.c_void_p()
        rc = self.ctx.L.mpg_pack_describe(self.h, C.byref(G), C.byref(ns), C.byref(tot), C.byref(po), C.byref(pi), C.byref(pv))
        if rc != 0:
            raise MpgError("mpg_pack_describe failed")
        self.ctx.sync()
        dt = np.float32 if self.sfx == "f32" else np.float64
        off = np.empty(ns.value + 1, np.int64)
        ind = np.empty(tot.value, np.int32)
        val = np.empty(tot.value, dt)
        for dst, src in ((off, po), (ind, pi), (val, pv)):
            if dst.nbytes:
                self.ctx._chk(self.ctx.L.mpg_memcpy_d2h(self.ctx.h, dst.ctypes.data_as(C.c_void_p), src, C.c_size_t(dst.nbytes)))
        return G.value, off, ind, val

    def rows(self):
        """(mode, lane_start, lane_len, lane_out, split_rows, chunk_base) of the plan as numpy arrays; mode 0 = plain slices of
        consecutive rows (the lane arrays are None), 1 = SELL-C-sigma (tests: bit-exact check against the oracle)"""
        import numpy as np
        mode, chunk, sigma, nl, ns, nc = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
        ps, pl, po, pr, pb = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        rc = self.ctx.L.mpg_pack_describe_rows(self.h, C.byref(mode), C.byref(chunk), C.byref(sigma), C.byref(nl), C.byref(ps), C.byref(pl), C.byref(po),
                                               C.byref(ns), C.byref(nc), C.byref(pr), C.byref(pb))
        if rc != 0:
            raise MpgError("mpg_pack_describe_rows failed")
        self.ctx.sync()
        if mode.value == 0:
            return 0, None, None, None, None, None

        def fetch(ptr, count):
            a = np.empty(count, np.int32)
            if count:
                self.ctx._chk(self.ctx.L.mpg_memcpy_d2h(self.ctx.h, a.ctypes.data_as(C.c_void_p), ptr, C.c_size_t(a.nbytes)))
            return a
        return mode.value, fetch(ps, nl.value), fetch(pl, nl.value), fetch(po, nl.value), fetch(pr, ns.value), fetch(pb, ns.value + 1)

    def __del__(self):
        try:
            if self.h:
                self.ctx.L.mpg_pack_destroy(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass


class IluJacobi:
    """mpg_ilu_jacobi handle: ILU_Jacobi<Type>(ilu, steps) of types.hpp:251-372; sfx 'f32' | 'f64' is Type"""

    def __init__(self, ctx, A, ilu_vals64, steps, sfx="f32"):
        self.ctx, self.A, self.sfx = ctx, A, sfx
        self.h = C.c_void_p()
        ctx._chk(getattr(ctx.L, "mpg_ilu_jacobi_create_" + sfx)(ctx.h, A.h, _ptr(ilu_vals64), C.c_int(steps), C.byref(self.h)))

    def apply(self, x):
        self.ctx._chk(getattr(self.ctx.L, "mpg_ilu_jacobi_apply_" + self.sfx)(self.ctx.h, self.h, _ptr(x)))

    def mv(self, lower, alpha, x, beta, y):
        self.ctx._chk(getattr(self.ctx.L, "mpg_ilu_jacobi_mv_" + self.sfx)(self.ctx.h, self.h, C.c_int(int(lower)), _sc(x, alpha), _ptr(x), _sc(x, beta), _ptr(y)))

    def __del__(self):
        try:
            if self.h:
                if self.ctx.h:          # the context may already be closed at interpreter exit: never hand the library a null context
                    self.ctx.sync()
                self.ctx.L.mpg_ilu_jacobi_destroy(self.h)
                self.h = C.c_void_p()
        except Exception:
            pass


class Context:
    def __init__(self, device=0):
        import torch
        self.L = load_library()
        if not torch.cuda.is_available():
            raise MpgError("no CUDA device: the B200 backend has no CPU fallback")
        self.device = device
        self.h = C.c_void_p()
        rc = self.L.mpg_ctx_create(C.c_int(device), C.byref(self.h))
        if rc != 0:
            raise MpgError(f"mpg_ctx_create failed with code {rc}")
        # run on torch's current stream so torch allocations / copies and our kernels are stream-ordered
        self.use_torch_stream()

    def use_torch_stream(self):
        import torch
        s = torch.cuda.current_stream(self.device).cuda_stream
        self._chk(self.L.mpg_ctx_set_stream(self.h, C.c_void_p(s)))

    def close(self):
        if self.h:
            self.L.mpg_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise MpgError(f"mpgmres_b200 error {rc}: {self.L.mpg_last_error(self.h).decode()}")

    def sync(self):
        if not self.h:
            raise MpgError("context is closed")
        self._chk(self.L.mpg_sync(self.h))

    def launches(self):
        return self.L.mpg_launch_count(self.h)

    def set_tuning(self, key, value):
        self._chk(self.L.mpg_set_tuning(self.h, key.encode(), C.c_int(value)))

    def get_tuning(self, key):
        v = C.c_int()
        self._chk(self.L.mpg_get_tuning(self.h, key.encode(), C.byref(v)))
        return v.value

    PROF_CLASSES = ["spmv_f32", "spmv_f64", "vpass", "gemvn", "elementwise", "reduce", "small", "gemvt"]

    def debug_timing(self, buf):
        self._chk(self.L.mpg_debug_timing(self.h, _ptr(buf) if buf is not None else None))

    def prof_enable(self, on=True):
        self._chk(self.L.mpg_prof_enable(self.h, C.c_int(int(on))))

    def prof_reset(self):
        self._chk(self.L.mpg_prof_reset(self.h))

    def prof_get(self):
        """{class: {"ms", "bytes", "launches"}} accumulated since the last reset (device time from CUDA events)"""
        out = {}
        for i, name in enumerate(self.PROF_CLASSES):
            ms, by, ln = C.c_double(), C.c_double(), C.c_int64()
            self._chk(self.L.mpg_prof_get(self.h, C.c_int(i), C.byref(ms), C.byref(by), C.byref(ln)))
            out[name] = {"ms": ms.value, "bytes": by.value, "launches": ln.value}
        return out

    # ---- generators ----
    def gen(self, spec):
        """'lap2d:N' | 'cd27:N' | 'powerlaw:n[:seed[:lmin[:gmax]]]' -> (row_map, inds, vals64) torch CUDA tensors"""
        import torch
        kind, *a = spec.split(":")
        a = [int(v) for v in a]
        dev = f"cuda:{self.device}"
        if kind in ("lap2d", "cd27"):
            N = a[0]
            n = N * N if kind == "lap2d" else N ** 3
            nnz = self.L.mpg_lap2d_nnz(N) if kind == "lap2d" else self.L.mpg_cd27_nnz(N)
            rm = torch.empty(n + 1, dtype=torch.int32, device=dev)
            ind = torch.empty(nnz, dtype=torch.int32, device=dev)
            val = torch.empty(nnz, dtype=torch.float64, device=dev)
            fn = self.L.mpg_gen_lap2d if kind == "lap2d" else self.L.mpg_gen_cd27
            self._chk(fn(self.h, C.c_int64(N), _ptr(rm), _ptr(ind), _ptr(val)))
            return rm, ind, val
        if kind == "powerlaw":
            n = a[0]
            seed = a[1] if len(a) > 1 else 7
            lmin = a[2] if len(a) > 2 else 2
            gmax = a[3] if len(a) > 3 else 15
            rm = torch.empty(n + 1, dtype=torch.int32, device=dev)
            nnz = C.c_int64()
            self._chk(self.L.mpg_gen_powerlaw_rowmap(self.h, C.c_int64(n), C.c_uint64(seed), C.c_int(lmin), C.c_int(gmax), _ptr(rm), C.byref(nnz)))
            ind = torch.empty(nnz.value, dtype=torch.int32, device=dev)
            val = torch.empty(nnz.value, dtype=torch.float64, device=dev)
            self._chk(self.L.mpg_gen_powerlaw_fill(self.h, C.c_int64(n), C.c_uint64(seed), C.c_int(lmin), C.c_int(gmax), _ptr(rm), _ptr(ind), _ptr(val)))
            return rm, ind, val
        raise ValueError(spec)

    GEN_KINDS = {"lap2d": 0, "cd27": 1, "powerlaw": 2}

    @classmethod
    def gen_params(cls, spec):
        """spec -> (kind, size, seed, lmin, gmax, n_rows)"""
        kind, *a = spec.split(":")
        a = [int(v) for v in a]
        size = a[0]
        seed = a[1] if len(a) > 1 else 7
        lmin = a[2] if len(a) > 2 else 2
        gmax = a[3] if len(a) > 3 else 15
        n = size * size if kind == "lap2d" else (size ** 3 if kind == "cd27" else size)
        return cls.GEN_KINDS[kind], size, seed, lmin, gmax, n

    def gen_rowmap(self, spec):
        """global row map of a synthetic matrix alone (int32 CUDA tensor, n + 1): what nnz-balanced split points need"""
        import torch
        k, size, seed, lmin, gmax, n = self.gen_params(spec)
        rm = torch.empty(n + 1, dtype=torch.int32, device=f"cuda:{self.device}")
        self._chk(self.L.mpg_gen_rowmap(self.h, C.c_int(k), C.c_int64(size), C.c_uint64(seed), C.c_int(lmin), C.c_int(gmax), _ptr(rm)))
        return rm

    def gen_slab(self, spec, lo, hi):
        """rows [lo, hi) of a synthetic matrix: (row_map_local, inds with GLOBAL columns, vals64); no rank builds the global matrix"""
        import torch
        k, size, seed, lmin, gmax, n = self.gen_params(spec)
        dev = f"cuda:{self.device}"
        rm = torch.empty(hi - lo + 1, dtype=torch.int32, device=dev)
        nnz = C.c_int64()
        self._chk(self.L.mpg_gen_slab_rowmap(self.h, C.c_int(k), C.c_int64(size), C.c_uint64(seed), C.c_int(lmin), C.c_int(gmax), C.c_int64(lo), C.c_int64(hi),
                                             _ptr(rm), C.byref(nnz)))
        ind = torch.empty(nnz.value, dtype=torch.int32, device=dev)
        val = torch.empty(nnz.value, dtype=torch.float64, device=dev)
        self._chk(self.L.mpg_gen_slab_fill(self.h, C.c_int(k), C.c_int64(size), C.c_uint64(seed), C.c_int(lmin), C.c_int(gmax), C.c_int64(lo), C.c_int64(hi),
                                           _ptr(rm), _ptr(ind), _ptr(val)))
        return rm, ind, val

    def rand_vect(self, n, seed=42):
        import numpy as np
        out = np.empty(n, dtype=np.float64)
        self._chk(self.L.mpg_rand_vect_host(C.c_int64(n), C.c_uint32(seed), _ptr(out)))
        return out

    # ---- BLAS-1 ----
    def dot(self, x, y):
        r = (C.c_float if _sfx(x) == "f32" else C.c_double)()
        self._chk(getattr(self.L, "mpg_dot_" + _sfx(x))(self.h, C.c_int64(x.numel()), _ptr(x), _ptr(y), C.byref(r)))
        return r.value

    def nrm2(self, x):
        r = (C.c_float if _sfx(x) == "f32" else C.c_double)()
        self._chk(getattr(self.L, "mpg_nrm2_" + _sfx(x))(self.h, C.c_int64(x.numel()), _ptr(x), C.byref(r)))
        return r.value

    def dot_dev(self, x, y, out):
        self._chk(getattr(self.L, "mpg_dot_dev_" + _sfx(x))(self.h, C.c_int64(x.numel()), _ptr(x), _ptr(y), _ptr(out)))

    def nrm2_dev(self, x, out):
        self._chk(getattr(self.L, "mpg_nrm2_dev_" + _sfx(x))(self.h, C.c_int64(x.numel()), _ptr(x), _ptr(out)))

    def axpy(self, alpha, x, y):
        self._chk(getattr(self.L, "mpg_axpy_" + _sfx(x))(self.h, C.c_int64(x.numel()), _sc(x, alpha), _ptr(x), _ptr(y)))

    def axpy_dev(self, alpha_dev, x, y):
        self._chk(getattr(self.L, "mpg_axpy_dev_" + _sfx(x))(self.h, C.c_int64(x.numel()), _ptr(alpha_dev), _ptr(x), _ptr(y)))

    def naxpy_dev(self, alpha_dev, x, y):
        self._chk(getattr(self.L, "mpg_naxpy_dev_" + _sfx(x))(self.h, C.c_int64(x.numel()), _ptr(alpha_dev), _ptr(x), _ptr(y)))

    def scal(self, alpha, x, y):
        self._chk(getattr(self.L, "mpg_scal_" + _sfx(x))(self.h, C.c_int64(x.numel()), _sc(x, alpha), _ptr(x), _ptr(y)))

    def scal_dev(self, alpha_dev, x, y):
        self._chk(getattr(self.L, "mpg_scal_dev_" + _sfx(x))(self.h, C.c_int64(x.numel()), _ptr(alpha_dev), _ptr(x), _ptr(y)))

    def copy(self, x, y):
        self._chk(getattr(self.L, f"mpg_copy_{_sfx(x)}_{_sfx(y)}")(self.h, C.c_int64(x.numel()), _ptr(x), _ptr(y)))

    def fill(self, alpha, x):
        self._chk(getattr(self.L, "mpg_fill_" + _sfx(x))(self.h, C.c_int64(x.numel()), _sc(x, alpha), _ptr(x)))

    def gdmv(self, alpha, diag, x, beta, y):
        self._chk(getattr(self.L, "mpg_gdmv_" + _sfx(x))(self.h, C.c_int64(x.numel()), _sc(x, alpha), _ptr(diag), _ptr(x), _sc(x, beta), _ptr(y)))

    # ---- Givens / LS ----
    def rotg(self, a, b, c, s):
        self._chk(getattr(self.L, "mpg_rotg_" + _sfx(a))(self.h, _ptr(a), _ptr(b), _ptr(c), _ptr(s)))

    def rot(self, a, b, c, s):
        self._chk(getattr(self.L, "mpg_rot_" + _sfx(a))(self.h, _ptr(a), _ptr(b), _ptr(c), _ptr(s)))

    def rot_vec(self, k, a, c, s):
        self._chk(getattr(self.L, "mpg_rot_vec_" + _sfx(a))(self.h, C.c_int64(k), _ptr(a), _ptr(c), _ptr(s)))

    def trsv(self, A, n, ld, x, upper=True, trans=False):
        self._chk(getattr(self.L, "mpg_trsv_" + _sfx(A))(self.h, C.c_int(int(upper)), C.c_int(int(trans)), C.c_int64(n), _ptr(A), C.c_int64(ld), _ptr(x)))

    def givens_step(self, k, h, ldh, cs, sn, s, resid):
        self._chk(getattr(self.L, "mpg_givens_step_" + _sfx(h))(self.h, C.c_int64(k), _ptr(h), C.c_int64(ldh), _ptr(cs), _ptr(sn), _ptr(s), _ptr(resid)))

    # ---- BLAS-2 / sparse ----
    def gemv(self, trans, nrows_base, ncols_base, alpha, M, ld, x, beta, y):
        self._chk(getattr(self.L, "mpg_gemv_" + _sfx(M))(self.h, C.c_int(int(trans)), C.c_int64(nrows_base), C.c_int64(ncols_base), _sc(M, alpha),
                                                         _ptr(M), C.c_int64(ld), _ptr(x), _sc(M, beta), _ptr(y)))

    def spmv(self, A, vals, alpha, x, beta, y):
        self._chk(getattr(self.L, "mpg_spmv_" + _sfx(vals))(self.h, A.h, _ptr(vals), _sc(vals, alpha), _ptr(x), _sc(vals, beta), _ptr(y)))

    def spmv_packed(self, P, alpha, x, beta, y):
        self._chk(getattr(self.L, "mpg_spmv_packed_" + P.sfx)(self.h, P.h, _sc(x, alpha), _ptr(x), _sc(x, beta), _ptr(y)))

    def spmv_jacobi(self, A, vals, diag, x, y):
        self._chk(getattr(self.L, "mpg_spmv_jacobi_" + _sfx(vals))(self.h, A.h, _ptr(vals), _ptr(diag), _ptr(x), _ptr(y)))

    def residual_cast(self, A, vals64, b, x, r64, w32):
        self._chk(self.L.mpg_residual_f64_cast_f32(self.h, A.h, _ptr(vals64), _ptr(b), _ptr(x), _ptr(r64), _ptr(w32)))

    def add_vector(self, orth, n, k, V, ldv, w, hcol):
        self._chk(getattr(self.L, "mpg_add_vector_" + _sfx(V))(self.h, C.c_int(ORTHS[orth]), C.c_int64(n), C.c_int64(k), _ptr(V), C.c_int64(ldv), _ptr(w), _ptr(hcol)))

    def jacobi_diag(self, A, vals, diag):
        self._chk(getattr(self.L, "mpg_jacobi_diag_" + _sfx(vals))(self.h, A.h, _ptr(vals), _ptr(diag)))

    # ---- ILU(0) + Jacobi sweeps ----
    def ilu0(self, A, vals64, eps_is_float=True):
        """fp64 ILU(0) factors on A's structure (new torch tensor)"""
        import torch
        out = torch.empty_like(vals64)
        self._chk(self.L.mpg_ilu0_f64(self.h, A.h, _ptr(vals64), C.c_int(int(eps_is_float)), _ptr(out)))
        return out

    def ilu0_levels(self, A):
        v = C.c_int()
        self._chk(self.L.mpg_ilu0_levels(self.h, A.h, C.byref(v)))
        return v.value

    # ---- solver ----
    @staticmethod
    def params(mode="mixed", orth="cgsr", conv="base", prec="identity", rlen=50, tol=1e-6, rtol=0.0, max_restarts=1000000, jacobi_steps=1):
        return GmresParams(MODES[mode], ORTHS[orth], CONVS[conv], PRECS[prec], rlen, tol, rtol, max_restarts, jacobi_steps)

    def gmres(self, A, vals64, b, x, vals32=None, hist_cap=None, **kw):
        """device-resident solve; x (torch float64 CUDA tensor) holds x0 on entry and the solution on exit"""
        import numpy as np
        p = self.params(**kw)
        st = GmresStats()
        cap_outer = min(int(p.max_restarts) + 2, 100000)
        cap_inner = hist_cap if hist_cap is not None else min(cap_outer * int(p.restart_length), 4000000)
        hi = np.zeros(max(cap_inner, 1), np.float64)
        ho = np.zeros(4 * cap_outer, np.float64)
        self._chk(self.L.mpg_gmres_solve(self.h, C.byref(p), A.h, _ptr(vals64), _ptr(vals32), _ptr(b), _ptr(x), C.byref(st),
                                         _ptr(hi), C.c_int64(cap_inner), _ptr(ho), C.c_int64(cap_outer)))
        return self._result(st, hi, ho, cap_inner, cap_outer)

    def gmres_host(self, row_map, inds, vals64, b, x, hist_cap=None, **kw):
        """end-to-end solve from HOST buffers (numpy arrays or pinned torch CPU tensors); x is updated in place"""
        import numpy as np
        p = self.params(**kw)
        st = GmresStats()
        cap_outer = min(int(p.max_restarts) + 2, 100000)
        cap_inner = hist_cap if hist_cap is not None else min(cap_outer * int(p.restart_length), 4000000)
        hi = np.zeros(max(cap_inner, 1), np.float64)
        ho = np.zeros(4 * cap_outer, np.float64)
        nrows = (row_map.numel() if hasattr(row_map, "numel") else len(row_map)) - 1
        nnz = inds.numel() if hasattr(inds, "numel") else len(inds)
        self._chk(self.L.mpg_gmres_solve_host(self.h, C.byref(p), C.c_int(nrows), C.c_int64(nnz), _ptr(row_map), _ptr(inds), _ptr(vals64),
                                              _ptr(b), _ptr(x), C.byref(st), _ptr(hi), C.c_int64(cap_inner), _ptr(ho), C.c_int64(cap_outer)))
        return self._result(st, hi, ho, cap_inner, cap_outer)

    @staticmethod
    def _result(st, hi, ho, cap_inner, cap_outer):
        res = {f: getattr(st, f) for f, _ in GmresStats._fields_}
        res["hist_inner"] = hi[:min(st.n_hist_inner, cap_inner)].copy()
        res["hist_outer"] = ho[:4 * min(st.n_hist_outer, cap_outer)].reshape(-1, 4).copy()
        return res
