// ORACLE — TEST INFRASTRUCTURE ONLY.
// cblas_* for the reference's kernels_mkl.cpp.  Level-1/2 routines that libtorch_cpu.so's embedded oneMKL exports
// through its Fortran interface are forwarded to it (genuine MKL arithmetic and threading); the four it does not
// export are restated from the netlib reference BLAS: ?nrm2 (scaled sum of squares), ?rotg, ?rot, ?trsv.
#include <cmath>

#include "mkl.h"

extern "C" {
float sdot_(const int*, const float*, const int*, const float*, const int*);
double ddot_(const int*, const double*, const int*, const double*, const int*);
void saxpy_(const int*, const float*, const float*, const int*, float*, const int*);
void daxpy_(const int*, const double*, const double*, const int*, double*, const int*);
void sscal_(const int*, const float*, float*, const int*);
void dscal_(const int*, const double*, double*, const int*);
void sgemv_(const char*, const int*, const int*, const float*, const float*, const int*, const float*, const int*, const float*, float*, const int*);
void dgemv_(const char*, const int*, const int*, const double*, const double*, const int*, const double*, const int*, const double*, double*, const int*);
}

namespace {
// netlib ?nrm2 (reference BLAS 3.8 form): scale / ssq recurrence
template <class T>
T nrm2_ref(int n, const T* x, int incx) {
    if (n < 1 || incx < 1) return T(0);
    if (n == 1) return std::fabs(x[0]);
    T scale = 0, ssq = 1;
    for (int i = 0; i < n; ++i) {
        const T v = x[(long)i * incx];
        if (v != T(0)) {
            const T a = std::fabs(v);
            if (scale < a) { ssq = T(1) + ssq * (scale / a) * (scale / a); scale = a; }
            else ssq += (a / scale) * (a / scale);
        }
    }
    return scale * std::sqrt(ssq);
}
template <class T>
void rotg_ref(T* a, T* b, T* c, T* s) {
    const T roe = (std::fabs(*a) > std::fabs(*b)) ? *a : *b;
    const T scale = std::fabs(*a) + std::fabs(*b);
    T r, z;
    if (scale == T(0)) { *c = 1; *s = 0; r = 0; z = 0; }
    else {
        const T as = *a / scale, bs = *b / scale;
        r = scale * std::sqrt(as * as + bs * bs);
        r = std::copysign(T(1), roe) * r;
        *c = *a / r;
        *s = *b / r;
        z = 1;
        if (std::fabs(*a) > std::fabs(*b)) z = *s;
        if (std::fabs(*b) >= std::fabs(*a) && *c != T(0)) z = T(1) / *c;
    }
    *a = r;
    *b = z;
}
template <class T>
void rot_ref(int n, T* x, int incx, T* y, int incy, T c, T s) {
    for (int i = 0; i < n; ++i) {
        T& xi = x[(long)i * incx];
        T& yi = y[(long)i * incy];
        const T t = c * xi + s * yi;
        yi = c * yi - s * xi;
        xi = t;
    }
}
// netlib ?trsv, column-major, incx = 1
template <class T>
void trsv_ref(CBLAS_UPLO uplo, CBLAS_TRANSPOSE trans, CBLAS_DIAG diag, int n, const T* A, int lda, T* x) {
    const bool nounit = diag == CblasNonUnit;
    if (trans == CblasNoTrans) {
        if (uplo == CblasUpper) {
            for (int j = n - 1; j >= 0; --j)
                if (x[j] != T(0)) {
                    if (nounit) x[j] /= A[j + (long)j * lda];
                    const T t = x[j];
                    for (int i = j - 1; i >= 0; --i) x[i] -= t * A[i + (long)j * lda];
                }
        } else {
            for (int j = 0; j < n; ++j)
                if (x[j] != T(0)) {
                    if (nounit) x[j] /= A[j + (long)j * lda];
                    const T t = x[j];
                    for (int i = j + 1; i < n; ++i) x[i] -= t * A[i + (long)j * lda];
                }
        }
    } else {
        if (uplo == CblasUpper) {
            for (int j = 0; j < n; ++j) {
                T t = x[j];
                for (int i = 0; i < j; ++i) t -= A[i + (long)j * lda] * x[i];
                if (nounit) t /= A[j + (long)j * lda];
                x[j] = t;
            }
        } else {
            for (int j = n - 1; j >= 0; --j) {
                T t = x[j];
                for (int i = n - 1; i > j; --i) t -= A[i + (long)j * lda] * x[i];
                if (nounit) t /= A[j + (long)j * lda];
                x[j] = t;
            }
        }
    }
}
}  // namespace

extern "C" {
float cblas_sdot(MKL_INT n, const float* x, MKL_INT incx, const float* y, MKL_INT incy) { return sdot_(&n, x, &incx, y, &incy); }
double cblas_ddot(MKL_INT n, const double* x, MKL_INT incx, const double* y, MKL_INT incy) { return ddot_(&n, x, &incx, y, &incy); }
float cblas_snrm2(MKL_INT n, const float* x, MKL_INT incx) { return nrm2_ref<float>(n, x, incx); }
double cblas_dnrm2(MKL_INT n, const double* x, MKL_INT incx) { return nrm2_ref<double>(n, x, incx); }
void cblas_saxpy(MKL_INT n, float a, const float* x, MKL_INT incx, float* y, MKL_INT incy) { saxpy_(&n, &a, x, &incx, y, &incy); }
void cblas_daxpy(MKL_INT n, double a, const double* x, MKL_INT incx, double* y, MKL_INT incy) { daxpy_(&n, &a, x, &incx, y, &incy); }
void cblas_sscal(MKL_INT n, float a, float* x, MKL_INT incx) { sscal_(&n, &a, x, &incx); }
void cblas_dscal(MKL_INT n, double a, double* x, MKL_INT incx) { dscal_(&n, &a, x, &incx); }
void cblas_srotg(float* a, float* b, float* c, float* s) { rotg_ref<float>(a, b, c, s); }
void cblas_drotg(double* a, double* b, double* c, double* s) { rotg_ref<double>(a, b, c, s); }
void cblas_srot(MKL_INT n, float* x, MKL_INT incx, float* y, MKL_INT incy, float c, float s) { rot_ref<float>(n, x, incx, y, incy, c, s); }
void cblas_drot(MKL_INT n, double* x, MKL_INT incx, double* y, MKL_INT incy, double c, double s) { rot_ref<double>(n, x, incx, y, incy, c, s); }
void cblas_sgemv(CBLAS_LAYOUT, CBLAS_TRANSPOSE t, MKL_INT m, MKL_INT n, float alpha, const float* a, MKL_INT lda, const float* x, MKL_INT incx, float beta,
                 float* y, MKL_INT incy) {
    const char tr = (t == CblasNoTrans) ? 'N' : 'T';
    sgemv_(&tr, &m, &n, &alpha, a, &lda, x, &incx, &beta, y, &incy);
}
void cblas_dgemv(CBLAS_LAYOUT, CBLAS_TRANSPOSE t, MKL_INT m, MKL_INT n, double alpha, const double* a, MKL_INT lda, const double* x, MKL_INT incx,
                 double beta, double* y, MKL_INT incy) {
    const char tr = (t == CblasNoTrans) ? 'N' : 'T';
    dgemv_(&tr, &m, &n, &alpha, a, &lda, x, &incx, &beta, y, &incy);
}
void cblas_strsv(CBLAS_LAYOUT, CBLAS_UPLO u, CBLAS_TRANSPOSE t, CBLAS_DIAG d, MKL_INT n, const float* a, MKL_INT lda, float* x, MKL_INT) {
    trsv_ref<float>(u, t, d, n, a, lda, x);
}
void cblas_dtrsv(CBLAS_LAYOUT, CBLAS_UPLO u, CBLAS_TRANSPOSE t, CBLAS_DIAG d, MKL_INT n, const double* a, MKL_INT lda, double* x, MKL_INT) {
    trsv_ref<double>(u, t, d, n, a, lda, x);
}
}
