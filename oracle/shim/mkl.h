/* ORACLE — TEST INFRASTRUCTURE ONLY.
 * CBLAS subset used by the reference's kernels_mkl.cpp.  libtorch_cpu.so exports MKL's Fortran-interface BLAS
 * (sdot_, ddot_, saxpy_, daxpy_, sscal_, dscal_, sgemv_, dgemv_) but not its CBLAS wrappers, nor ?nrm2/?rotg/?rot/?trsv
 * (SURVEY.md §8c).  oracle/shim/mkl_shim.cpp forwards the former to genuine MKL and restates the latter from netlib. */
#ifndef ORACLE_SHIM_MKL_H
#define ORACLE_SHIM_MKL_H
#include "mkl_spblas.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_LAYOUT;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
typedef enum { CblasUpper = 121, CblasLower = 122 } CBLAS_UPLO;
typedef enum { CblasNonUnit = 131, CblasUnit = 132 } CBLAS_DIAG;
float cblas_sdot(MKL_INT n, const float* x, MKL_INT incx, const float* y, MKL_INT incy);
double cblas_ddot(MKL_INT n, const double* x, MKL_INT incx, const double* y, MKL_INT incy);
float cblas_snrm2(MKL_INT n, const float* x, MKL_INT incx);
double cblas_dnrm2(MKL_INT n, const double* x, MKL_INT incx);
void cblas_saxpy(MKL_INT n, float a, const float* x, MKL_INT incx, float* y, MKL_INT incy);
void cblas_daxpy(MKL_INT n, double a, const double* x, MKL_INT incx, double* y, MKL_INT incy);
void cblas_sscal(MKL_INT n, float a, float* x, MKL_INT incx);
void cblas_dscal(MKL_INT n, double a, double* x, MKL_INT incx);
void cblas_srotg(float* a, float* b, float* c, float* s);
void cblas_drotg(double* a, double* b, double* c, double* s);
void cblas_srot(MKL_INT n, float* x, MKL_INT incx, float* y, MKL_INT incy, float c, float s);
void cblas_drot(MKL_INT n, double* x, MKL_INT incx, double* y, MKL_INT incy, double c, double s);
void cblas_sgemv(CBLAS_LAYOUT l, CBLAS_TRANSPOSE t, MKL_INT m, MKL_INT n, float alpha, const float* a, MKL_INT lda, const float* x, MKL_INT incx,
                 float beta, float* y, MKL_INT incy);
void cblas_dgemv(CBLAS_LAYOUT l, CBLAS_TRANSPOSE t, MKL_INT m, MKL_INT n, double alpha, const double* a, MKL_INT lda, const double* x, MKL_INT incx,
                 double beta, double* y, MKL_INT incy);
void cblas_strsv(CBLAS_LAYOUT l, CBLAS_UPLO u, CBLAS_TRANSPOSE t, CBLAS_DIAG d, MKL_INT n, const float* a, MKL_INT lda, float* x, MKL_INT incx);
void cblas_dtrsv(CBLAS_LAYOUT l, CBLAS_UPLO u, CBLAS_TRANSPOSE t, CBLAS_DIAG d, MKL_INT n, const double* a, MKL_INT lda, double* x, MKL_INT incx);
#ifdef __cplusplus
}
#endif
#endif
