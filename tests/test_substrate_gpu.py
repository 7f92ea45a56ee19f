"""include/b200/substrate.hpp — the Kokkos-free Scalar / Vect / MultiVect / SparseMatrix handles + operator surface:
tests/cpp/substrate_test.cpp drives gemv / trsv / rot / rotg / dot / naxpy / spmv through sub-range, sub-block and transpose-flag
views exactly as Orthogonalization.hpp:38,48,58,63,69,83-84,121-123 and gmres.cpp:219-222,276-303 do and checks every step
against host loops.  It is compiled with plain g++ (no nvcc, no Kokkos) against the C ABI."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "substrate_test.cpp")
LIBDIR = os.path.join(ROOT, "icl-mixed-precision-gmres_b200", "lib")


def build(tmp_path):
    exe = str(tmp_path / "substrate_test")
    cmd = ["g++", "-std=c++14", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe, "-L", LIBDIR, "-lmpgmres_b200",
           f"-Wl,-rpath,{LIBDIR}"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-4000:]
    assert "warning" not in out.stderr, out.stderr[-4000:]
    return exe


def _env():
    import torch   # the shared library resolves libcudart through the copy torch ships
    tl = os.path.join(os.path.dirname(torch.__file__), "lib")
    extra = [tl] + [os.path.join(r, d) for r in (os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), "nvidia"),) if os.path.isdir(r)
                    for d in ("cuda_runtime/lib",) if os.path.isdir(os.path.join(r, d))]
    return dict(os.environ, LD_LIBRARY_PATH=":".join(extra + [os.environ.get("LD_LIBRARY_PATH", "")]))


def test_substrate_compiles_without_kokkos_or_nvcc(tmp_path):
    """CPU side: one header + the C ABI is all a C++ caller needs (no Kokkos, no CUDA headers)"""
    exe = build(tmp_path)
    src = open(os.path.join(ROOT, "include", "b200", "substrate.hpp")).read()
    assert "Kokkos_Core" not in src and "cuda_runtime" not in src
    assert os.path.exists(exe)


@pytest.mark.gpu
def test_substrate_drives_the_operator_surface_through_subviews(tmp_path):
    exe = build(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=_env())
    assert out.returncode == 0, (out.stdout[-4000:], out.stderr[-2000:])
    assert "substrate ok" in out.stdout
