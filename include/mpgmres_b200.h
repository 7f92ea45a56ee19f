/*
 * mpgmres_b200.h — C ABI of the B200-native (sm_100a) backend for the mixed-precision GMRES hot path of
 * iamsonderr/icl-mixed-precision-gmres.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch/Kokkos types.  Every entry point
 * names the reference interface it replaces (file:line relative to the reference tree).  The reference
 * resolves its operator surface (kernels.hpp) at link time to one translation unit per backend
 * (kernels_cuda.cpp / kernels_mkl.cpp); a new backend = template specialisations that forward to these
 * functions (see INTEGRATION.md and include/b200/kernels_b200.hpp).
 *
 * Conventions
 *   - All vector / matrix pointers are DEVICE pointers unless the name says _host.
 *   - Dense matrices are column-major with leading dimension `ld` (types.hpp:115-118, LayoutLeft).
 *   - CSR is 0-based int32 row_map[nrows+1] / inds[nnz] (types_cuda.hpp:66-70); fp32 and fp64 value arrays
 *     share one structure (types_cuda.hpp:82-91), which is why mpg_csr carries no values.
 *   - Work is stream-ordered on the context's stream and asynchronous unless the function returns a value
 *     to the host (the reference relies on the same semantics).
 *   - Return value: 0 on success, non-zero error code otherwise; mpg_last_error() gives the message.  The
 *     reference surface returns void and ignores library status codes; the C++ shim aborts on non-zero.
 *   - There is NO CPU fallback: without a CUDA device every compute entry point fails with MPG_ERR_CUDA.
 */
#ifndef MPGMRES_B200_H
#define MPGMRES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPG_OK 0
#define MPG_ERR_CUDA 1
#define MPG_ERR_ARG 2
#define MPG_ERR_NCCL 3
#define MPG_ERR_STATE 4

typedef struct mpg_ctx mpg_ctx; /* replaces CudaLibSingleton, types_cuda.hpp:9-36: stream, reduction scratch, comm */
typedef struct mpg_csr mpg_csr; /* CSR structure + SpMV plan; replaces cusparseMatDescr_t, types_cuda.hpp:50-60 */

/* ---- context ----------------------------------------------------------------------------------------- */
const char* mpg_version(void);
int mpg_ctx_create(int device, mpg_ctx** out);
int mpg_ctx_destroy(mpg_ctx* ctx);
/* A new context runs on its own non-blocking stream.  set_stream takes any cudaStream_t; NULL is the CUDA legacy
 * default stream, which is what the reference runs everything on. */
int mpg_ctx_set_stream(mpg_ctx* ctx, void* cuda_stream);
int mpg_ctx_use_own_stream(mpg_ctx* ctx);
void* mpg_ctx_stream(mpg_ctx* ctx);
int mpg_sync(mpg_ctx* ctx); /* Device::execution_space().fence(), gmres.cpp:113,225 */
const char* mpg_last_error(mpg_ctx* ctx);
int mpg_num_sms(mpg_ctx* ctx);
int64_t mpg_launch_count(mpg_ctx* ctx);              /* kernels launched through this context so far */
int mpg_set_tuning(mpg_ctx* ctx, const char* key, int value); /* kernel variant knobs, see DESIGN.md */
int mpg_get_tuning(mpg_ctx* ctx, const char* key, int* value);
/* Development aid: with a device buffer of >= 8 * 160 uint64 attached, every CTA of the staged V-pass kernels stores
 * %globaltimer at its phase boundaries (start, first tile in, last tile done, partials written, finish done);
 * NULL detaches (tools/vpass_timeline.py). */
int mpg_debug_timing(mpg_ctx* ctx, unsigned long long* device_buf);

/* ---- per-kernel-class device timers (CUDA events on the launching stream; the reference has none) -- */
enum { MPG_PROF_SPMV_F32 = 0, MPG_PROF_SPMV_F64 = 1, MPG_PROF_VPASS = 2, MPG_PROF_GEMVN = 3, MPG_PROF_ELEMENTWISE = 4,
       MPG_PROF_REDUCE = 5, MPG_PROF_SMALL = 6, MPG_PROF_GEMVT = 7 };
int mpg_prof_enable(mpg_ctx* ctx, int on);
int mpg_prof_reset(mpg_ctx* ctx);
/* synchronises the stream, folds pending measurements; returns accumulated device ms, algorithmic bytes, launches */
int mpg_prof_get(mpg_ctx* ctx, int cls, double* ms, double* bytes, int64_t* launches);

/* ---- device memory (Kokkos::View allocation / deep_copy, types.hpp:18,60,118) -------------------------- */
int mpg_malloc(mpg_ctx* ctx, size_t bytes, void** dptr); /* zero-filled like a Kokkos::View */
int mpg_free(mpg_ctx* ctx, void* dptr);
int mpg_memcpy_h2d(mpg_ctx* ctx, void* dst, const void* src_host, size_t bytes);
int mpg_memcpy_d2h(mpg_ctx* ctx, void* dst_host, const void* src, size_t bytes); /* synchronises */
int mpg_memcpy_d2d(mpg_ctx* ctx, void* dst, const void* src, size_t bytes);
int mpg_memset_zero(mpg_ctx* ctx, void* dst, size_t bytes);

/* ---- BLAS-1 --------------------------------------------------------------------------------------------
 * dot   kernels.hpp:33-37   (kernels_cuda.cpp:111-165)     nrm2  kernels.hpp:40-44  (kernels_cuda.cpp:167-209)
 * axpy  kernels.hpp:47-51   (kernels_cuda.cpp:212-262)     naxpy kernels.hpp:59-60  (kernels_cuda.cpp:264-288)
 * scal  kernels.hpp:62-85   (kernels_cuda.cpp:291-392)     copy  kernels.hpp:11-30  fill kernels.hpp:88-101
 * gdmv  kernels.hpp:131-151
 * nrm2 accumulates the squares in double: the fp32 form is safe over the whole fp32 range without scaling (the role of BLAS
 * ?nrm2's scale/ssq recurrence, kernels_mkl.cpp:97-115); the fp64 form is the plain sum of squares, safe for |x| in
 * [1e-150, 1e150].
 * `_dev` variants take/return the scalar in device memory (the Scalar<T,Device> overloads); the others
 * take host scalars / return to the host (and therefore synchronise), like the reference's two forms. */
int mpg_dot_f32(mpg_ctx*, int64_t n, const float* x, const float* y, float* result_host);
int mpg_dot_f64(mpg_ctx*, int64_t n, const double* x, const double* y, double* result_host);
int mpg_dot_dev_f32(mpg_ctx*, int64_t n, const float* x, const float* y, float* result_dev);
int mpg_dot_dev_f64(mpg_ctx*, int64_t n, const double* x, const double* y, double* result_dev);
int mpg_nrm2_f32(mpg_ctx*, int64_t n, const float* x, float* result_host);
int mpg_nrm2_f64(mpg_ctx*, int64_t n, const double* x, double* result_host);
int mpg_nrm2_dev_f32(mpg_ctx*, int64_t n, const float* x, float* result_dev);
int mpg_nrm2_dev_f64(mpg_ctx*, int64_t n, const double* x, double* result_dev);
int mpg_axpy_f32(mpg_ctx*, int64_t n, float alpha, const float* x, float* y);
int mpg_axpy_f64(mpg_ctx*, int64_t n, double alpha, const double* x, double* y);
int mpg_axpy_dev_f32(mpg_ctx*, int64_t n, const float* alpha_dev, const float* x, float* y);
int mpg_axpy_dev_f64(mpg_ctx*, int64_t n, const double* alpha_dev, const double* x, double* y);
int mpg_naxpy_dev_f32(mpg_ctx*, int64_t n, const float* alpha_dev, const float* x, float* y);
int mpg_naxpy_dev_f64(mpg_ctx*, int64_t n, const double* alpha_dev, const double* x, double* y);
int mpg_scal_f32(mpg_ctx*, int64_t n, float alpha, const float* x, float* y); /* y = alpha*x; x may equal y */
int mpg_scal_f64(mpg_ctx*, int64_t n, double alpha, const double* x, double* y);
int mpg_scal_dev_f32(mpg_ctx*, int64_t n, const float* alpha_dev, const float* x, float* y);
int mpg_scal_dev_f64(mpg_ctx*, int64_t n, const double* alpha_dev, const double* x, double* y);
int mpg_copy_f32_f32(mpg_ctx*, int64_t n, const float* x, float* y);
int mpg_copy_f64_f64(mpg_ctx*, int64_t n, const double* x, double* y);
int mpg_copy_f64_f32(mpg_ctx*, int64_t n, const double* x, float* y); /* RN cast, gmres.cpp:163,175 */
int mpg_copy_f32_f64(mpg_ctx*, int64_t n, const float* x, double* y); /* Orthogonalization.hpp:71 */
int mpg_fill_f32(mpg_ctx*, int64_t n, float alpha, float* x);
int mpg_fill_f64(mpg_ctx*, int64_t n, double alpha, double* x);
int mpg_gdmv_f32(mpg_ctx*, int64_t n, float alpha, const float* diag, const float* x, float beta, float* y);
int mpg_gdmv_f64(mpg_ctx*, int64_t n, double alpha, const double* diag, const double* x, double beta, double* y);

/* ---- Givens / least squares (all scalars in device memory) -----------------------------------------------
 * rotg kernels.hpp:104-106 (kernels_cuda.cpp:394-420: BLAS rotg, then b := 0)
 * rot  kernels.hpp:109-114 (kernels_cuda.cpp:422-494)      trsv kernels.hpp:128-129 (kernels_cuda.cpp:538-572) */
int mpg_rotg_f32(mpg_ctx*, float* a, float* b, float* c, float* s);
int mpg_rotg_f64(mpg_ctx*, double* a, double* b, double* c, double* s);
int mpg_rot_f32(mpg_ctx*, float* a, float* b, const float* c, const float* s);
int mpg_rot_f64(mpg_ctx*, double* a, double* b, const double* c, const double* s);
int mpg_rot_vec_f32(mpg_ctx*, int64_t k, float* a, const float* c, const float* s); /* touches a[0..k] */
int mpg_rot_vec_f64(mpg_ctx*, int64_t k, double* a, const double* c, const double* s);
int mpg_trsv_f32(mpg_ctx*, int upper, int trans, int64_t n, const float* A, int64_t ld, float* x);
int mpg_trsv_f64(mpg_ctx*, int upper, int trans, int64_t n, const double* A, int64_t ld, double* x);
/* Fused: apply the k stored rotations to column k of H, generate rotation k, rotate s, store |s[k+1]| (as a
 * double) to resid_dev.  Replaces gmres.cpp:219-226 (4-5 launches + 1 blocking read) with one launch. */
int mpg_givens_step_f32(mpg_ctx*, int64_t k, float* h, int64_t ldh, float* cs, float* sn, float* s, double* resid_dev);
int mpg_givens_step_f64(mpg_ctx*, int64_t k, double* h, int64_t ldh, double* cs, double* sn, double* s, double* resid_dev);

/* ---- BLAS-2: gemv  kernels.hpp:118-125 (kernels_cuda.cpp:499-535) ----------------------------------------
 * y = alpha*op(M)*x + beta*y; M is nrows_base x ncols_base (the reference passes base dims + lda). */
int mpg_gemv_f32(mpg_ctx*, int trans, int64_t nrows_base, int64_t ncols_base, float alpha, const float* M, int64_t ld,
                 const float* x, float beta, float* y);
int mpg_gemv_f64(mpg_ctx*, int trans, int64_t nrows_base, int64_t ncols_base, double alpha, const double* M, int64_t ld,
                 const double* x, double beta, double* y);

/* ---- Sparse: spmv  kernels.hpp:159-160 (kernels_cuda.cpp:576-614) ----------------------------------------- */
int mpg_csr_create(mpg_ctx*, int nrows, int ncols, int64_t nnz, const int* row_map, const int* inds, mpg_csr** out);
int mpg_csr_destroy(mpg_csr* A);
int mpg_spmv_f32(mpg_ctx*, const mpg_csr* A, const float* vals, float alpha, const float* x, float beta, float* y);
int mpg_spmv_f64(mpg_ctx*, const mpg_csr* A, const double* vals, double alpha, const double* x, double beta, double* y);
/* Jacobi-preconditioned operator in one kernel: y = diag .* (A*x), i.e. spmv (gmres.cpp:100,212) followed by
 * Jacobi::apply = gdmv(1, diag, y, 0, y) (types.hpp:444-446), bit-identical to the two separate calls. */
int mpg_spmv_jacobi_f32(mpg_ctx*, const mpg_csr* A, const float* vals, const float* diag, const float* x, float* y);
int mpg_spmv_jacobi_f64(mpg_ctx*, const mpg_csr* A, const double* vals, const double* diag, const double* x, double* y);
/* Packed operator (sell.cu): the same matrix re-laid as 32-lane slices stored position-major (sliced ELLPACK), the
 * analogue of the per-matrix library handles the reference builds once per SparseMatrix (create_cuda_handles,
 * types_cuda.hpp:53-60).  One lane per row: the gathers of x coalesce for stencil-like matrices.  Matrices with uneven
 * row lengths (power-law) take the SELL-C-sigma form: rows longer than 256 nonzeros are cut into pieces, the pieces
 * are sorted by length inside windows of 4096, a fix-up kernel adds the pieces of each cut row in order; y stays in
 * caller order and results are bit-reproducible.  fp32 and fp64 value arrays share ONE packed index array.
 * mpg_pack_create returns MPG_OK with *out == NULL when even that form pads > 25 %; keep using mpg_spmv_* then.
 * The packed object refers to A (destroy it before A); mpg_pack_update re-packs the values after the caller changed
 * them.  mpg_gmres_solve packs the matrices it multiplies with itself. */
typedef struct mpg_packed mpg_packed;
int mpg_pack_create_f32(mpg_ctx*, const mpg_csr* A, const float* vals, mpg_packed** out);
int mpg_pack_create_f64(mpg_ctx*, const mpg_csr* A, const double* vals, mpg_packed** out);
int mpg_pack_update_f32(mpg_ctx*, mpg_packed* P, const float* vals);
int mpg_pack_update_f64(mpg_ctx*, mpg_packed* P, const double* vals);
int mpg_pack_destroy(mpg_packed* P);
/* Layout access for tests / diagnostics: group size (4), slice count, padded element count and the DEVICE arrays
 * slice_off[nslices + 1], inds[total], vals[total] (owned by the library; layout in csrc/sell.cu and DESIGN.md §2). */
int mpg_pack_describe(const mpg_packed* P, int* group, int* nslices, int64_t* total, const int64_t** slice_off, const int** inds,
                      const void** vals);
/* mode 0: lane = row (the lane arrays are NULL).  mode 1 (SELL-C-sigma): per lane, in sorted order, the CSR offset and
 * length of its piece and its destination (row index, or -1 - piece number for a piece of a cut row); the cut rows and
 * the first piece number of each (chunk_base[nsplit] = number of pieces).  DEVICE arrays owned by the library. */
int mpg_pack_describe_rows(const mpg_packed* P, int* mode, int* chunk, int* sigma, int* nlanes, const int** lane_start, const int** lane_len,
                           const int** lane_out, int* nsplit, int* nchunks, const int** split_rows, const int** chunk_base);
int mpg_spmv_packed_f32(mpg_ctx*, const mpg_packed* P, float alpha, const float* x, float beta, float* y);
int mpg_spmv_packed_f64(mpg_ctx*, const mpg_packed* P, double alpha, const double* x, double beta, double* y);
/* Fused outer residual, replaces gmres.cpp:173-175 (copy + fp64 SpMV + cast kernel):
 * r = b - A*x in fp64; w32 = (float) r; r64 may be NULL (then r is never stored). */
int mpg_residual_f64_cast_f32(mpg_ctx*, const mpg_csr* A, const double* vals, const double* b, const double* x,
                              double* r64, float* w32);

/* ---- ILU(0) + Jacobi-sweep triangular solves: ilu0 kernels.hpp:163-164 (kernels_cuda.cpp:714-791), ILU_Jacobi
 * types.hpp:251-372, ilu_jacobi_mv / ilusv_jacobi kernels.hpp:172-248.  (The exact triangular solves of ILU<>::apply -
 * cusparse csrsv2, kernels_cuda.cpp:617-695, removed from CUDA 12 - are not provided.)
 * mpg_ilu0_f64 factors the fp64 matrix on its own sparsity pattern (level-scheduled IKJ, bit-identical to the sequential
 * loop of kernels_mkl.cpp:451-484 with the diagonal positions filled in); eps_is_float selects the pivot-boost threshold
 * eps(Type) * max_i sum_j |a_ij| of ilu0<float> / ilu0<double>.  vals_out may not alias vals_in. */
typedef struct mpg_ilu_jacobi mpg_ilu_jacobi;
int mpg_ilu0_f64(mpg_ctx*, const mpg_csr* A, const double* vals_in, int eps_is_float, double* vals_out);
int mpg_ilu0_levels(mpg_ctx*, const mpg_csr* A, int* nlevels);   /* depth of the dependency schedule (diagnostics) */
/* ILU_Jacobi<Type>(ilu, steps): converts the factors to Type, extracts 1/diag (types.hpp:290-304) */
int mpg_ilu_jacobi_create_f32(mpg_ctx*, const mpg_csr* A, const double* ilu_vals, int steps, mpg_ilu_jacobi** out);
int mpg_ilu_jacobi_create_f64(mpg_ctx*, const mpg_csr* A, const double* ilu_vals, int steps, mpg_ilu_jacobi** out);
int mpg_ilu_jacobi_destroy(mpg_ilu_jacobi* M);
/* ilusv_jacobi: x <- approximately U^-1 L^-1 x with `steps` Jacobi sweeps per triangle, one fused launch per sweep */
int mpg_ilu_jacobi_apply_f32(mpg_ctx*, mpg_ilu_jacobi* M, float* x);
int mpg_ilu_jacobi_apply_f64(mpg_ctx*, mpg_ilu_jacobi* M, double* x);
/* ilu_jacobi_mv: lower: y = beta*y + alpha*(x + L x); upper: y = y - U x (alpha, beta ignored as in kernels.hpp:205-216) */
int mpg_ilu_jacobi_mv_f32(mpg_ctx*, const mpg_ilu_jacobi* M, int lower, float alpha, const float* x, float beta, float* y);
int mpg_ilu_jacobi_mv_f64(mpg_ctx*, const mpg_ilu_jacobi* M, int lower, double alpha, const double* x, double beta, double* y);

/* ---- Fused Arnoldi step: GS::add_vector, Orthogonalization.hpp:51-60 + the kernels at :76-136 ------------
 * orth: 0 CGS, 1 MGS, 2 CGSR<2> (CGS2).  V is n x (k+2) column-major with leading dimension ldv.
 * In:  w (A*v_k), V[:,0:k+1].  Out: hcol[0..k] coefficients, hcol[k+1] = ||w_orth||, V[:,k+1] = w_orth/hcol[k+1];
 * w is overwritten with w_orth.  No host synchronisation (the reference reads h(k+1,k) back, :56). */
int mpg_add_vector_f32(mpg_ctx*, int orth, int64_t n, int64_t k, float* V, int64_t ldv, float* w, float* hcol);
int mpg_add_vector_f64(mpg_ctx*, int orth, int64_t n, int64_t k, double* V, int64_t ldv, double* w, double* hcol);

/* ---- Solver: gmres.hpp:15-32 (gmres.cpp:24-245) behind one entry point ------------------------------------ */
enum { MPG_MODE_MIXED = 0, MPG_MODE_BASELINE = 1, MPG_MODE_SINGLE_PREC = 2, MPG_MODE_SINGLE = 3 }; /* gmres_perf_test.cpp:31-36 */
enum { MPG_ORTH_CGS = 0, MPG_ORTH_MGS = 1, MPG_ORTH_CGSR = 2 };                                    /* :17-22 */
enum { MPG_CONV_BASE = 0, MPG_CONV_RELPRECRES = 1, MPG_CONV_REPEAT = 2, MPG_CONV_ORTHLOSS = 3 };   /* :185-196 */
enum { MPG_PREC_IDENTITY = 0, MPG_PREC_JACOBI = 1, MPG_PREC_ILU_JACOBI = 2 };                      /* :24-29 (exact ilu: not provided) */

typedef struct mpg_gmres_params {
    int32_t mode, orth, conv, prec;
    int64_t restart_length; /* --rlen; 1 .. 255 (the fused passes keep one accumulator set per basis column: wider bases are rejected with MPG_ERR_ARG; the reference has no limit) */
    double tol;             /* --tol   */
    double restart_tol;     /* --rtol  */
    int64_t max_restarts;   /* --max-restarts */
    int64_t jacobi_steps;   /* --jacobi-steps (MPG_PREC_ILU_JACOBI), gmres_perf_test.cpp:323,386-387 */
} mpg_gmres_params;

typedef struct mpg_gmres_stats {
    int64_t status; /* 1 converged, 3 aborted (iteration_action, IterUtil.hpp:10-15) */
    int64_t total_iters;
    int64_t total_restarts;
    int64_t outer_i;
    double rel_prec_res; /* the printed "rel prec res norm", gmres.cpp:186 */
    double b_norm, Minvb_norm, A_norm;
    int64_t n_hist_inner, n_hist_outer;
    double solve_ms;   /* device time of the solve (CUDA events), excluding transfers */
    double h2d_ms, d2h_ms; /* only set by the _host entry point; h2d_ms = until the solve could start */
    int64_t launches;  /* kernels launched by this solve */
    /* _host entry point: time until the LAST input byte had landed (overlapped shape: inside the first restart cycle), bytes that
     * crossed the link host->device (the overlapped shape also sends the fp32 values the host cores cast), and the number of host
     * cast threads (0: serial shape) */
    double h2d_all_ms;
    int64_t h2d_bytes;
    int64_t host_overlap;
} mpg_gmres_stats;

/* Device-resident operands.  vals32 may be NULL (cast from vals64 internally, types_cuda.hpp:82-101).
 * hist_inner_host[cap_inner]: |s(k+1)|/Minvb_norm per inner iteration; hist_outer_host[4*cap_outer]:
 * {r_norm, b_norm + A_norm*x_norm, beta, x_norm} per check_initial.  Either may be NULL. */
int mpg_gmres_solve(mpg_ctx*, const mpg_gmres_params* p, const mpg_csr* A, const double* vals64, const float* vals32,
                    const double* b, double* x, mpg_gmres_stats* stats, double* hist_inner_host, int64_t cap_inner,
                    double* hist_outer_host, int64_t cap_outer);
/* End-to-end: HOST CSR + b in, x out (H2D, plan, solve, D2H inside).  x_host holds x0 on entry.  Host buffers should be pinned.
 * Mixed precision (knob host_overlap, default on): the indices go first while host threads cast the values to fp32 into a pinned
 * staging buffer, the solve starts when the fp32 operator is complete and the fp64 values land during the first restart cycle -
 * same bits as the serial shape (csrc/hostpath.cu). */
int mpg_gmres_solve_host(mpg_ctx*, const mpg_gmres_params* p, int nrows, int64_t nnz, const int* row_map_host,
                         const int* inds_host, const double* vals64_host, const double* b_host, double* x_host,
                         mpg_gmres_stats* stats, double* hist_inner_host, int64_t cap_inner, double* hist_outer_host,
                         int64_t cap_outer);
/* Jacobi<T>::get_diag_vals, types.hpp:395-430 */
int mpg_jacobi_diag_f32(mpg_ctx*, const mpg_csr* A, const float* vals, float* diag);
int mpg_jacobi_diag_f64(mpg_ctx*, const mpg_csr* A, const double* vals, double* diag);

/* ---- Synthetic inputs (the reference has none).  Definitions frozen in oracle/oracle.cpp. -- */
int64_t mpg_lap2d_nnz(int64_t N);
int64_t mpg_cd27_nnz(int64_t N);
int mpg_gen_lap2d(mpg_ctx*, int64_t N, int* row_map, int* inds, double* vals);
int mpg_gen_cd27(mpg_ctx*, int64_t N, int* row_map, int* inds, double* vals);
int mpg_gen_powerlaw_rowmap(mpg_ctx*, int64_t n, uint64_t seed, int lmin, int gmax, int* row_map, int64_t* nnz_host);
int mpg_gen_powerlaw_fill(mpg_ctx*, int64_t n, uint64_t seed, int lmin, int gmax, const int* row_map, int* inds, double* vals);
/* Row-range forms: rows [lo, hi) of the same matrices with GLOBAL column indices and a local row map, so that every rank of a
 * multi-GPU run generates only its slab (mpg_dist_setup renumbers the columns).  kind: 0 lap2d (size = N), 1 cd27 (size = N),
 * 2 powerlaw (size = n; seed, lmin, gmax as above).  mpg_gen_rowmap: the global row map alone (nnz-balanced split points). */
int mpg_gen_slab_rowmap(mpg_ctx*, int kind, int64_t size, uint64_t seed, int lmin, int gmax, int64_t lo, int64_t hi, int* row_map_local,
                        int64_t* nnz_local_host);
int mpg_gen_slab_fill(mpg_ctx*, int kind, int64_t size, uint64_t seed, int lmin, int gmax, int64_t lo, int64_t hi, int* row_map_local,
                      int* inds_global, double* vals);
int mpg_gen_rowmap(mpg_ctx*, int kind, int64_t size, uint64_t seed, int lmin, int gmax, int* row_map);
/* gmres_perf_test.cpp:39-51 rand_vect (host; libstdc++ mt19937 + uniform_real_distribution<float>) */
int mpg_rand_vect_host(int64_t n, uint32_t seed, double* out_host);

/* ---- multi-GPU: one process per GPU, 1-D row blocks (new functionality) --------------------------
 * NCCL is loaded at run time.  Rank 0 makes a 128-byte unique id, the host plumbing (torch.distributed) broadcasts it,
 * every rank creates its communicator, describes its halo plan and attaches it to its context.  From then on the
 * reductions behind dot / nrm2 / gemv-T / add_vector are all-reduced over the ranks, and mpg_halo_exchange fills the
 * halo tail of an SpMV input; mpg_gmres_solve takes the LOCAL slab (nrows = n_local, ncols = n_local + n_halo). */
typedef struct mpg_dist mpg_dist;
int mpg_nccl_unique_id(void* id128_host);
int mpg_dist_create(mpg_ctx*, const void* id128_host, int rank, int world, mpg_dist** out);
int mpg_dist_destroy(mpg_dist* d);
/* peers: for each neighbour, the ascending local row indices it needs from us (device array, not owned) and the
 * [offset, offset+count) range of our halo that it owns. */
int mpg_dist_set_partition(mpg_ctx*, mpg_dist* d, int64_t n_global, int64_t n_local, int64_t n_halo, int npeers, const int* peer_ranks_host,
                           const int64_t* send_counts_host, const int* const* send_idx_dev_ptrs_host, const int64_t* recv_offsets_host,
                           const int64_t* recv_counts_host);
/* Peer-memory mailboxes for the in-kernel all-reduce: every rank exports a 64-byte CUDA IPC handle, the host plumbing
 * all-gathers them (rank order) and every rank maps its peers.  Without this step reductions go through ncclAllReduce. */
int mpg_dist_mailbox_handle(mpg_ctx*, mpg_dist* d, void* handle64_host);
int mpg_dist_open_mailboxes(mpg_ctx*, mpg_dist* d, const void* handles_world_x_64_host);
/* Peer-memory halo inboxes (after mpg_dist_set_partition).  remote_offsets / remote_nhalo: for each neighbour of the plan,
 * where this rank's rows land in that neighbour's halo and that neighbour's halo length.  Without this step the halo goes
 * through pack + ncclSend/ncclRecv. */
int mpg_dist_halo_handle(mpg_ctx*, mpg_dist* d, void* handle64_host);
int mpg_dist_open_halo(mpg_ctx*, mpg_dist* d, const void* handles_world_x_64_host, const int64_t* remote_offsets_host,
                       const int64_t* remote_nhalo_host);
/* Native set-up from this rank's slab ALONE (no rank ever holds the global matrix): bounds_host[world + 1] = first row of every
 * rank (mpg_partition_bounds / mpg_partition_bounds_nnz), inds_dev = the slab's column indices, GLOBAL on entry, renumbered
 * [local | halo] on return.  Builds halo and send lists on the device, exchanges them and the CUDA IPC handles of mailboxes
 * and inboxes over the communicator; equivalent to mpg_dist_set_partition + mailbox / halo hand-shakes.  Collective. */
int mpg_dist_setup(mpg_ctx*, mpg_dist* d, int64_t n_global, const int64_t* bounds_host, int64_t nnz_local, int* inds_dev, int64_t* n_halo);
int mpg_dist_halo_cols(mpg_ctx*, const mpg_dist* d, int64_t* halo_cols_host /* n_halo */);
/* The device part of mpg_dist_setup alone, without a communicator: what rank `rank` of P computes for its slab - halo columns
 * (ascending), renumbered inds (in place), per-owner offsets into the halo (P + 1) and the owner-local row index of every halo
 * slot (the lists it would request).  halo_cols_dev / need_idx_dev: capacity `cap` ints (may be NULL to query n_halo only -
 * inds is renumbered either way).  One GPU can so check every rank's plan against the oracle. */
int mpg_partition_slab_dev(mpg_ctx*, int64_t n_global, int P, const int64_t* bounds_host, int rank, int64_t nnz_local, int* inds_dev,
                           int64_t* n_halo, int* halo_cols_dev, int* need_idx_dev, int64_t cap, int64_t* owner_off_host);
int mpg_dist_peer_info(mpg_ctx*, const mpg_dist* d, int i, int* peer_rank, int64_t* send_count, int64_t* recv_offset, int64_t* recv_count,
                       int* send_idx_host /* may be NULL */);
int mpg_ctx_attach_dist(mpg_ctx*, mpg_dist* d); /* NULL detaches */
int mpg_dist_info(const mpg_dist* d, int* rank, int* world, int64_t* n_global, int64_t* n_local, int64_t* n_halo);
int mpg_halo_exchange_f32(mpg_ctx*, float* x_ext);  /* x_ext = [n_local owned | n_halo halo slots] */
int mpg_halo_exchange_f64(mpg_ctx*, double* x_ext);
int mpg_allreduce_sum_f64(mpg_ctx*, double* buf_dev, int64_t count);

/* ---- MatrixMarket ingest: LoadMatrix<S>(), LoadMatrix.hpp:17-154, into canonical CSR (host).  Arrays are malloc'ed by
 * the library and released with mpg_host_free.  errbuf receives the reference's exception text on failure. ------------ */
int mpg_mm_read_host(const char* path, int* nrows, int* ncols, int64_t* nnz, int** row_map_host, int** inds_host, double** vals_host,
                     char* errbuf, int errlen);
/* LoadVector<S>(file, col), LoadMatrix.hpp:156-233 (--bpath, gmres_perf_test.cpp:417-421): column `col` of an array or
 * coordinate MatrixMarket file as n doubles (malloc'ed, release with mpg_host_free) */
int mpg_mm_read_vector_host(const char* path, int col, int64_t* n, double** vals_host, char* errbuf, int errlen);
/* Partition-aware ingest: rows [lo, hi) of the canonical CSR of a MatrixMarket file, read WITHOUT ever holding the other
 * rows - what one rank of a multi-GPU run ingests (hi < 0: to the last row).  The file is streamed in blocks; memory = O(entries of the
 * slab).  Local row map (row_map_local[0] = 0), GLOBAL column indices (mpg_dist_setup renumbers them), nnz of the slab and of the whole
 * matrix.  row_map_global_host may be NULL; otherwise it receives the global row map (n + 1 ints, for mpg_partition_bounds_nnz; lo = hi = 0
 * reads nothing else).  Equal to rows [lo, hi) of mpg_mm_read_host bit for bit; same error texts. */
int mpg_mm_read_slab_host(const char* path, int64_t lo, int64_t hi, int* nrows, int* ncols, int64_t* nnz_global, int64_t* nnz_local,
                          int** row_map_local_host, int** inds_host, double** vals_host, int** row_map_global_host, char* errbuf, int errlen);
void mpg_host_free(void* p);

/* ---- 1-D row partition (new functionality, host-side, bit-exact vs the oracle) ------------- */
int mpg_partition_bounds(int64_t n, int P, int64_t* bounds_host /* P+1 */);
/* nnz-balanced split points: bounds[k] = the smallest row i with row_map[i] >= floor(k * nnz / P); bounds[0] = 0,
 * bounds[P] = n.  row_map_host is the GLOBAL row map (4 (n + 1) bytes - small even when the matrix itself does not fit). */
int mpg_partition_bounds_nnz(int64_t n, int P, const int* row_map_host, int64_t* bounds_host /* P+1 */);
/* halo_cols_host/local_inds_host may be NULL to query sizes.  Returns the halo count through *n_halo. */
int mpg_partition_local(int64_t n, int P, int r, const int* row_map_host, const int* inds_host, int64_t* n_halo,
                        int64_t* halo_cols_host, int* local_inds_host);

#ifdef __cplusplus
}
#endif
#endif /* MPGMRES_B200_H */
