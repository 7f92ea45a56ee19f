"""CPU checks of the drop-in boundary: the C-ABI library loads, exports exactly what include/mpgmres_b200.h declares,
refuses to compute without a GPU (no CPU fallback), and its host-only helpers agree with the oracle bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest


def test_library_builds_and_exports_every_declared_symbol(g):
    import importlib.util
    spec = importlib.util.spec_from_file_location("mpg_build", os.path.join(g.PACKAGE_DIR, "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    lib = b.build()
    assert os.path.exists(lib)
    declared, exported = set(g.header_symbols()), set(g.exported_symbols())
    assert declared, "header parse found nothing"
    assert declared - exported == set(), f"declared but not exported: {sorted(declared - exported)}"
    assert exported - declared == set(), f"exported but not declared: {sorted(exported - declared)}"
    L = g.load_library()
    assert b"sm_100a" in L.mpg_version()


def test_binary_is_sm100a_and_uses_tma(g):
    out = subprocess.run(["cuobjdump", "-lelf", g.library_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out
    sass = subprocess.run(["cuobjdump", "-sass", g.library_path()], capture_output=True, text=True).stdout
    assert "UTMALDG" in sass, "the V-pass kernel must stage its tiles with TMA tensor loads"
    assert "SYNCS" in sass


def test_no_cpu_fallback(g):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = g.load_library()
    h = C.c_void_p()
    assert L.mpg_ctx_create(C.c_int(0), C.byref(h)) != 0  # fails loudly, no silent host path
    with pytest.raises(g.MpgError):
        g.Context(0)


def test_host_helpers_match_oracle_bit_exactly(g, orc):
    L = g.load_library()
    out = np.empty(1000, np.float64)
    assert L.mpg_rand_vect_host(C.c_int64(1000), C.c_uint32(42), out.ctypes.data_as(C.c_void_p)) == 0
    np.testing.assert_array_equal(out, orc.rand_vect(1000, 42))
    for spec in ["lap2d:9", "cd27:6", "powerlaw:400"]:
        rm, ind, val = orc.gen(spec)
        n = len(rm) - 1
        for P in (1, 2, 3, 8):
            b = np.empty(P + 1, np.int64)
            assert L.mpg_partition_bounds(C.c_int64(n), C.c_int(P), b.ctypes.data_as(C.c_void_p)) == 0
            np.testing.assert_array_equal(b, orc.partition_bounds(n, P))
            for r in range(P):
                halo_o, li_o = orc.partition_local(n, P, r, rm, ind)
                nh = C.c_int64()
                assert L.mpg_partition_local(C.c_int64(n), C.c_int(P), C.c_int(r), rm.ctypes.data_as(C.c_void_p),
                                             ind.ctypes.data_as(C.c_void_p), C.byref(nh), None, None) == 0
                assert nh.value == len(halo_o)
                halo = np.empty(nh.value, np.int64)
                li = np.empty(len(li_o), np.int32)
                assert L.mpg_partition_local(C.c_int64(n), C.c_int(P), C.c_int(r), rm.ctypes.data_as(C.c_void_p),
                                             ind.ctypes.data_as(C.c_void_p), C.byref(nh), halo.ctypes.data_as(C.c_void_p),
                                             li.ctypes.data_as(C.c_void_p)) == 0
                np.testing.assert_array_equal(halo, halo_o)   # halo index sets: bit-exact
                np.testing.assert_array_equal(li, li_o)       # remapped local column indices: bit-exact


def test_size_queries(g, orc):
    L = g.load_library()
    for N in (1, 2, 7, 256):
        assert L.mpg_lap2d_nnz(N) == 5 * N * N - 4 * N
        assert L.mpg_cd27_nnz(N) == (3 * N - 2) ** 3
    assert L.mpg_cd27_nnz(256) == 449455096  # config 3


def test_nnz_balanced_split_points_match_oracle(g, orc):
    """mpg_partition_bounds_nnz vs the oracle's linear-scan definition, bit-exact; each slab's nonzero count is
    within one row of nnz / P"""
    for spec in ["lap2d:9", "cd27:6", "powerlaw:400", "powerlaw:20000"]:
        rm, ind, val = orc.gen(spec)
        n, nnz = len(rm) - 1, int(rm[-1])
        for P in (1, 2, 3, 4, 8):
            b = np.array(g.dist.bounds_nnz(rm, P), np.int64)
            np.testing.assert_array_equal(b, orc.partition_bounds_nnz(rm, P))
            assert b[0] == 0 and b[-1] == n and np.all(np.diff(b) >= 0)
            share = np.diff(rm[b].astype(np.int64))
            assert np.all(np.abs(share - nnz / P) <= np.diff(rm).max() + 1)
    # the torch restatement of the plan honours explicit split points (bit-exact vs the oracle's range form)
    import torch
    rm, ind, val = orc.gen("powerlaw:3000")
    n = len(rm) - 1
    b = g.dist.bounds_nnz(rm, 3)
    for r in range(3):
        rml, li, v, halo = g.dist.local_slab(torch.from_numpy(rm), torch.from_numpy(ind), torch.from_numpy(val), n, r, 3, b)
        halo_o, li_o = orc.partition_local_range(b[r], b[r + 1], rm, ind)
        np.testing.assert_array_equal(halo.numpy(), halo_o)
        np.testing.assert_array_equal(li.numpy(), li_o)
