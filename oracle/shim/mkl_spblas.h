/* ORACLE — TEST INFRASTRUCTURE ONLY.
 * Hand-declared prototypes of the Intel oneMKL Inspector-Executor sparse BLAS entry points the reference uses
 * (types_mkl.hpp, kernels_mkl.cpp).  The MKL SDK is not installed in this image, but torch/lib/libtorch_cpu.so
 * statically embeds oneMKL and exports these symbols; oracle/ref.mk links against that library,
 * so the arithmetic behind mkl_sparse_?_mv is Intel's, not a restatement.  Enumerator values follow the public
 * mkl_spblas.h. */
#ifndef ORACLE_SHIM_MKL_SPBLAS_H
#define ORACLE_SHIM_MKL_SPBLAS_H
#ifdef __cplusplus
extern "C" {
#endif
typedef int MKL_INT;
typedef enum { SPARSE_STATUS_SUCCESS = 0, SPARSE_STATUS_NOT_INITIALIZED = 1, SPARSE_STATUS_ALLOC_FAILED = 2, SPARSE_STATUS_INVALID_VALUE = 3,
               SPARSE_STATUS_EXECUTION_FAILED = 4, SPARSE_STATUS_INTERNAL_ERROR = 5, SPARSE_STATUS_NOT_SUPPORTED = 6 } sparse_status_t;
typedef enum { SPARSE_OPERATION_NON_TRANSPOSE = 10, SPARSE_OPERATION_TRANSPOSE = 11, SPARSE_OPERATION_CONJUGATE_TRANSPOSE = 12 } sparse_operation_t;
typedef enum { SPARSE_MATRIX_TYPE_GENERAL = 20, SPARSE_MATRIX_TYPE_SYMMETRIC = 21, SPARSE_MATRIX_TYPE_HERMITIAN = 22, SPARSE_MATRIX_TYPE_TRIANGULAR = 23,
               SPARSE_MATRIX_TYPE_DIAGONAL = 24, SPARSE_MATRIX_TYPE_BLOCK_TRIANGULAR = 25, SPARSE_MATRIX_TYPE_BLOCK_DIAGONAL = 26 } sparse_matrix_type_t;
typedef enum { SPARSE_INDEX_BASE_ZERO = 0, SPARSE_INDEX_BASE_ONE = 1 } sparse_index_base_t;
typedef enum { SPARSE_FILL_MODE_LOWER = 40, SPARSE_FILL_MODE_UPPER = 41, SPARSE_FILL_MODE_FULL = 42 } sparse_fill_mode_t;
typedef enum { SPARSE_DIAG_NON_UNIT = 50, SPARSE_DIAG_UNIT = 51 } sparse_diag_type_t;
struct matrix_descr { sparse_matrix_type_t type; sparse_fill_mode_t mode; sparse_diag_type_t diag; };
struct sparse_matrix;
typedef struct sparse_matrix* sparse_matrix_t;

sparse_status_t mkl_sparse_s_create_csr(sparse_matrix_t* A, sparse_index_base_t indexing, MKL_INT rows, MKL_INT cols, MKL_INT* rows_start,
                                        MKL_INT* rows_end, MKL_INT* col_indx, float* values);
sparse_status_t mkl_sparse_d_create_csr(sparse_matrix_t* A, sparse_index_base_t indexing, MKL_INT rows, MKL_INT cols, MKL_INT* rows_start,
                                        MKL_INT* rows_end, MKL_INT* col_indx, double* values);
sparse_status_t mkl_sparse_destroy(sparse_matrix_t A);
sparse_status_t mkl_sparse_s_mv(sparse_operation_t op, float alpha, const sparse_matrix_t A, struct matrix_descr descr, const float* x, float beta, float* y);
sparse_status_t mkl_sparse_d_mv(sparse_operation_t op, double alpha, const sparse_matrix_t A, struct matrix_descr descr, const double* x, double beta, double* y);
sparse_status_t mkl_sparse_s_trsv(sparse_operation_t op, float alpha, const sparse_matrix_t A, struct matrix_descr descr, const float* x, float* y);
sparse_status_t mkl_sparse_d_trsv(sparse_operation_t op, double alpha, const sparse_matrix_t A, struct matrix_descr descr, const double* x, double* y);
int MKL_Get_Max_Threads(void);
#ifdef __cplusplus
}
#endif
#endif
