"""helpers shared by the parity tests"""
import numpy as np

EPS = {np.dtype(np.float32): 2.0 ** -24, np.dtype(np.float64): 2.0 ** -53}  # unit roundoff


def dev(a, device="cuda:0"):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(device)


def host(t):
    return t.detach().cpu().numpy()


def fortran_dev(M, device="cuda:0"):
    """numpy (n, k) array -> torch CUDA tensor holding the column-major buffer (k*n elements)"""
    import torch
    return torch.from_numpy(np.asfortranarray(M).ravel(order="F").copy()).to(device)


def summation_bound(terms_abs_sum, nterms, dtype, c=8.0):
    """any-order recursive summation of nterms products in `dtype`: |err| <= ~nterms*u*sum|terms| worst case;
    blocked/pairwise implementations sit near sqrt(nterms)*u.  c*sqrt(n)*u*sum|terms| + a few ulps."""
    u = EPS[np.dtype(dtype)]
    return c * np.sqrt(max(nterms, 1)) * u * terms_abs_sum + 4 * u * terms_abs_sum


def problem(orc, spec, seed=42, bscale=1.0):
    """(row_map, inds, vals64, x_true, b) with b = A x_true in fp64 (gmres_perf_test.cpp:413-416); bscale multiplies x_true
    and b afterwards (scaled-magnitude cases of the golden file)"""
    rm, ind, val = orc.gen(spec)
    n = len(rm) - 1
    xt = orc.rand_vect(n, seed)
    b = np.zeros(n)
    orc.spmv(rm, ind, val, 1.0, xt, 0.0, b)
    if bscale != 1.0:
        xt, b = xt * bscale, b * bscale
    return rm, ind, val, xt, b
