// kernels_b200.cpp — the operator surface of kernels.hpp specialised for the B200 device: the counterpart of the
// reference's kernels_cuda.cpp (cuBLAS / legacy cuSPARSE / Kokkos lambdas), forwarding every call to the C ABI of
// libmpgmres_b200.so.  Compile it into the reference build next to kernels_mkl.cpp.
// Each specialisation names the kernels_cuda.cpp lines it stands in for.
#include "kernels.hpp"
#include "types_b200.hpp"

#define CTX B200Backend::context()

// ---- BLAS-1 ------------------------------------------------------------------------------------------------------
#define B200_BLAS1(T, SFX)                                                                                                        \
    template <> T dot<T, B200>(Vect<T, B200> x, Vect<T, B200> y) { /* kernels_cuda.cpp:111-137 */                                 \
        assert(x.n() == y.n());                                                                                                   \
        T r;                                                                                                                      \
        B200_CHECK(mpg_dot_##SFX(CTX, x.n(), x.data(), y.data(), &r));                                                            \
        return r;                                                                                                                 \
    }                                                                                                                             \
    template <> void dot<T, B200>(Vect<T, B200> x, Vect<T, B200> y, Scalar<T, B200> result) { /* :139-165 */                      \
        assert(x.n() == y.n());                                                                                                   \
        B200_CHECK(mpg_dot_dev_##SFX(CTX, x.n(), x.data(), y.data(), result.data()));                                             \
    }                                                                                                                             \
    template <> T nrm2<T, B200>(Vect<T, B200> x) { /* :167-187 */                                                                 \
        T r;                                                                                                                      \
        B200_CHECK(mpg_nrm2_##SFX(CTX, x.n(), x.data(), &r));                                                                     \
        return r;                                                                                                                 \
    }                                                                                                                             \
    template <> void nrm2<T, B200>(Vect<T, B200> x, Scalar<T, B200> result) { /* :189-209 */                                      \
        B200_CHECK(mpg_nrm2_dev_##SFX(CTX, x.n(), x.data(), result.data()));                                                      \
    }                                                                                                                             \
    template <> void axpy<T, B200>(T alpha, Vect<T, B200> x, Vect<T, B200> y) { /* :212-234 */                                    \
        assert(x.n() == y.n());                                                                                                   \
        B200_CHECK(mpg_axpy_##SFX(CTX, x.n(), alpha, x.data(), y.data()));                                                        \
    }                                                                                                                             \
    template <> void axpy<T, B200>(Scalar<T, B200> alpha, Vect<T, B200> x, Vect<T, B200> y) { /* :236-262 */                      \
        assert(x.n() == y.n());                                                                                                   \
        B200_CHECK(mpg_axpy_dev_##SFX(CTX, x.n(), alpha.data(), x.data(), y.data()));                                             \
    }                                                                                                                             \
    template <> void naxpy<T, B200>(Scalar<T, B200> alpha, Vect<T, B200> x, Vect<T, B200> y) { /* :264-288 */                     \
        assert(x.n() == y.n());                                                                                                   \
        B200_CHECK(mpg_naxpy_dev_##SFX(CTX, x.n(), alpha.data(), x.data(), y.data()));                                            \
    }                                                                                                                             \
    template <> void scal<T, B200>(T alpha, Vect<T, B200> x) { /* :291-307 */                                                     \
        B200_CHECK(mpg_scal_##SFX(CTX, x.n(), alpha, x.data(), x.data()));                                                        \
    }                                                                                                                             \
    template <> void scal<T, B200>(T alpha, Vect<T, B200> x, Vect<T, B200> y) { /* :309-331, one pass instead of copy + scal */   \
        assert(x.n() == y.n());                                                                                                   \
        B200_CHECK(mpg_scal_##SFX(CTX, x.n(), alpha, x.data(), y.data()));                                                        \
    }                                                                                                                             \
    template <> void scal<T, B200>(Scalar<T, B200> alpha, Vect<T, B200> x, Vect<T, B200> y) { /* :333-359 */                      \
        assert(x.n() == y.n());                                                                                                   \
        B200_CHECK(mpg_scal_dev_##SFX(CTX, x.n(), alpha.data(), x.data(), y.data()));                                             \
    }                                                                                                                             \
    template <> void scal<T, B200>(T alpha, Scalar<T, B200> x, Scalar<T, B200> y) { /* :362-392 */                                \
        B200_CHECK(mpg_scal_##SFX(CTX, 1, alpha, x.data(), y.data()));                                                            \
    }                                                                                                                             \
    template <> void scal<T, B200>(Scalar<T, B200> alpha, Scalar<T, B200> x, Scalar<T, B200> y) { /* kernels.hpp:84-85 */         \
        B200_CHECK(mpg_scal_dev_##SFX(CTX, 1, alpha.data(), x.data(), y.data()));                                                 \
    }                                                                                                                             \
    template <> void rotg<T, B200>(Scalar<T, B200> a, Scalar<T, B200> b, Scalar<T, B200> c, Scalar<T, B200> s) { /* :394-420 */   \
        B200_CHECK(mpg_rotg_##SFX(CTX, a.data(), b.data(), c.data(), s.data()));                                                  \
    }                                                                                                                             \
    template <> void rot<T, B200>(Scalar<T, B200> a, Scalar<T, B200> b, Scalar<T, B200> c, Scalar<T, B200> s) { /* :422-446 */    \
        B200_CHECK(mpg_rot_##SFX(CTX, a.data(), b.data(), c.data(), s.data()));                                                   \
    }                                                                                                                             \
    template <> void rot<T, B200>(Vect<T, B200> a, Vect<T, B200> c, Vect<T, B200> s) { /* :448-494: k = c.n() rotations */        \
        B200_CHECK(mpg_rot_vec_##SFX(CTX, c.n(), a.data(), c.data(), s.data()));                                                  \
    }                                                                                                                             \
    template <> void gemv<T, B200>(T alpha, MultiVect<T, B200> matrix, Vect<T, B200> x, T beta, Vect<T, B200> y) { /* :499-535 */ \
        assert(matrix.ncols() == x.n());                                                                                          \
        assert(matrix.nrows() == y.n());                                                                                          \
        B200_CHECK(mpg_gemv_##SFX(CTX, matrix.transposed() ? 1 : 0, matrix.nrows_base(), matrix.ncols_base(), alpha, matrix.data(), \
                                  matrix.stride(), x.data(), beta, y.data()));                                                    \
    }                                                                                                                             \
    template <> void trsv<T, B200>(const char* upper, MultiVect<T, B200> matrix, Vect<T, B200> x) { /* :538-572 */                \
        assert(matrix.ncols() == matrix.nrows());                                                                                 \
        assert(matrix.ncols() == x.n());                                                                                          \
        B200_CHECK(mpg_trsv_##SFX(CTX, 'U' == *upper, matrix.transposed() ? 1 : 0, matrix.nrows(), matrix.data(), matrix.stride(), x.data())); \
    }                                                                                                                             \
    template <> void spmv<T, B200>(T alpha, SparseMatrix<T, B200> matrix, Vect<T, B200> x, T beta, Vect<T, B200> y) { /* :576-614 */ \
        assert(matrix.ncols() == x.n());                                                                                          \
        assert(matrix.nrows() == y.n());                                                                                          \
        if (matrix.is_transposed()) Kokkos::abort("transposed SpMV (condest.cpp only) is not provided by the B200 backend\n");    \
        if (matrix.packed()) B200_CHECK(mpg_spmv_packed_##SFX(CTX, matrix.packed(), alpha, x.data(), beta, y.data()));            \
        else B200_CHECK(mpg_spmv_##SFX(CTX, matrix.plan(), matrix.vals_data(), alpha, x.data(), beta, y.data()));                 \
    }                                                                                                                             \
    template <> void gdmv<T, B200>(T alpha, Vect<T, B200> diag, Vect<T, B200> x, T beta, Vect<T, B200> y) { /* kernels.hpp:131-146 */ \
        B200_CHECK(mpg_gdmv_##SFX(CTX, diag.n(), alpha, diag.data(), x.data(), beta, y.data()));                                  \
    }                                                                                                                             \
    template <> void ilusv<T, B200>(ILU<T, B200>, Vect<T, B200>) { /* :617-695 csrsv2: not provided */                            \
        Kokkos::abort("ilusv (exact triangular solves) is not provided by the B200 backend; use --prec ilu_jacobi\n");           \
    }                                                                                                                             \
    template <> ILU<T, B200> ilu0<T, B200>(SparseMatrix<double, B200> matrix) { /* :714-791 csrilu02 -> level-scheduled IKJ */    \
        assert(matrix.nrows() == matrix.ncols());                                                                                 \
        Kokkos::View<double*, typename B200::memory_space> vals("ilu::vals64", (size_t)matrix.nnz());                             \
        B200_CHECK(mpg_ilu0_f64(CTX, matrix.plan(), matrix.vals_data(), sizeof(T) == 4, vals.data()));                            \
        Kokkos::View<T*, typename B200::memory_space> vals_type("ilu::vals", (size_t)matrix.nnz());   /* type_convert, :697-712 */ \
        copy(Vect<double, B200>(vals), Vect<T, B200>(vals_type));                                                                 \
        B200_CHECK(mpg_sync(CTX));                                                                                                \
        return ILU<T, B200>(matrix.nrows(), matrix.nnz(), matrix.row_map_, matrix.inds_, vals_type);                              \
    }                                                                                                                             \
    template <> void ilusv_jacobi<T, B200>(ILU_Jacobi<T, B200> ilu, Vect<T, B200> x) { /* kernels.hpp:227-248, one launch per sweep */ \
        assert(ilu.n() == (int)x.n());                                                                                            \
        B200_CHECK(mpg_ilu_jacobi_apply_##SFX(CTX, ilu.handles().h.get(), x.data()));                                             \
    }                                                                                                                             \
    template <> void ilu_jacobi_mv<T, B200>(bool lower, T alpha, ILU_Jacobi<T, B200> ilu, Vect<T, B200> x, T beta, Vect<T, B200> y) { /* kernels.hpp:172-216 */ \
        B200_CHECK(mpg_ilu_jacobi_mv_##SFX(CTX, ilu.handles().h.get(), lower ? 1 : 0, alpha, x.data(), beta, y.data()));          \
    }

B200_BLAS1(float, f32)
B200_BLAS1(double, f64)

// ---- element-wise assign with conversion: the fp64 <-> fp32 casts (kernels.hpp:11-20) ------------------------------------
template <> void copy<double, float, B200>(Vect<double, B200> x, Vect<float, B200> y) { assert(x.n() == y.n()); B200_CHECK(mpg_copy_f64_f32(CTX, x.n(), x.data(), y.data())); }
template <> void copy<float, double, B200>(Vect<float, B200> x, Vect<double, B200> y) { assert(x.n() == y.n()); B200_CHECK(mpg_copy_f32_f64(CTX, x.n(), x.data(), y.data())); }
template <> void copy<float, float, B200>(Vect<float, B200> x, Vect<float, B200> y) { assert(x.n() == y.n()); B200_CHECK(mpg_copy_f32_f32(CTX, x.n(), x.data(), y.data())); }
template <> void copy<double, double, B200>(Vect<double, B200> x, Vect<double, B200> y) { assert(x.n() == y.n()); B200_CHECK(mpg_copy_f64_f64(CTX, x.n(), x.data(), y.data())); }

// ---- GS::add_vector (Orthogonalization.hpp:51-60) fused on the device: orth 0 CGS, 1 MGS, 2 CGSR<2> ----------------------
// h(0:k+1, k) is contiguous (column k of the column-major Hessenberg array); V(:,k+1) is written by the library.
namespace Orthogonalization {
typedef CGS_Kernel<float, B200> B200_CGS_f;
typedef MGS_Kernel<float, B200> B200_MGS_f;
typedef CGSR_Kernel<float, B200, 2> B200_CGS2_f;
typedef CGS_Kernel<double, B200> B200_CGS_d;
typedef MGS_Kernel<double, B200> B200_MGS_d;
typedef CGSR_Kernel<double, B200, 2> B200_CGS2_d;
#define B200_ADD_VECTOR(T, SFX, KERNEL, ORTH)                                                                                              \
    template <> void GS<T, KERNEL, B200>::add_vector(const size_t k, Vect<T, B200> w, MultiVect<T, B200> h) {                              \
        T* hcol = h.data() + k * h.stride();                                                                                               \
        B200_CHECK(mpg_add_vector_##SFX(CTX, ORTH, (int64_t)v.nrows_base(), (int64_t)k, v.data(), (int64_t)v.stride(), w.data(), hcol));   \
    }
B200_ADD_VECTOR(float, f32, B200_CGS_f, MPG_ORTH_CGS)
B200_ADD_VECTOR(float, f32, B200_MGS_f, MPG_ORTH_MGS)
B200_ADD_VECTOR(float, f32, B200_CGS2_f, MPG_ORTH_CGSR)
B200_ADD_VECTOR(double, f64, B200_CGS_d, MPG_ORTH_CGS)
B200_ADD_VECTOR(double, f64, B200_MGS_d, MPG_ORTH_MGS)
B200_ADD_VECTOR(double, f64, B200_CGS2_d, MPG_ORTH_CGSR)
#undef B200_ADD_VECTOR
}  // namespace Orthogonalization
