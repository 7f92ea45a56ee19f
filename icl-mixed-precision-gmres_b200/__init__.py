"""icl-mixed-precision-gmres_b200 — B200-native (sm_100a) backend for the mixed-precision GMRES hot path.

The product is the C-ABI shared library (csrc/ -> lib/libmpgmres_b200.so, header include/mpgmres_b200.h) and
the C++ drop-in headers under include/b200/.  This Python package is plumbing for tests and bench.py only:
a ctypes binding that passes torch CUDA tensor pointers through the C ABI.  There is no CPU fallback:
importing works anywhere, but every compute call needs the CUDA library and a CUDA device.
"""
from .binding import (Context, CSR, Packed, IluJacobi, GmresParams, GmresStats, MODES, ORTHS, CONVS, PRECS, load_library, library_path,
                      exported_symbols, header_symbols, MpgError, read_matrix_market, read_matrix_market_vector,
                      read_matrix_market_slab)

from . import dist  # noqa: E402,F401  (multi-GPU plumbing: partition builder + DistContext)

__all__ = ["dist", "Context", "CSR", "Packed", "IluJacobi", "GmresParams", "GmresStats", "MODES", "ORTHS", "CONVS", "PRECS", "load_library",
           "library_path", "exported_symbols", "header_symbols", "MpgError", "read_matrix_market", "read_matrix_market_vector", "read_matrix_market_slab"]
