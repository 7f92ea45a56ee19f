// ilu.cu — ILU(0) factorisation and the Jacobi-sweep application of its factors.
// Reference surface: ilu0<Type,Device> kernels.hpp:163-164 (Cuda: cusparseDcsrilu02 with level policy + numeric boost,
// kernels_cuda.cpp:714-791; MKL: the sequential IKJ loop of kernels_mkl.cpp:420-487), ILU_Jacobi<Type,Device> types.hpp:251-372,
// ilu_jacobi_mv / ilusv_jacobi kernels.hpp:172-248.  The exact triangular solves (ILU<>::apply -> csrsv2, kernels_cuda.cpp:617-695)
// stay out of scope: latency-bound, and the API no longer exists in CUDA 12.
//
// Factorisation: level-scheduled, one warp per row, fp64, IKJ on the sparsity pattern of A.
//   * level(i) = 1 + max level(k) over the lower-triangle columns k of row i, found by in-place relaxation sweeps (levels only
//     grow, every sweep is a pure function of the previous state or better, the fixed point is unique);
//   * rows are bucketed by level (order inside a level is irrelevant: rows of one level do not depend on each other);
//   * one launch per level; the warp of row i walks its lower entries k in ascending order: factor = a_ik / a_kk, then
//     a_ij = fma(-factor, a_kj, a_ij) for every j > k that both rows store (lanes take the entries of row k, binary search in
//     row i).  Every a_ij receives its updates in ascending k exactly like the sequential loop: results are bit-identical to
//     the restated IKJ factorisation in oracle/oracle.cpp, whatever the schedule.
//   * pivot boost as in the reference: |a_ii| < alpha = eps(Type) * max_i sum_j |a_ij|  ->  a_ii = +-alpha (rows i >= 1).
// The plan (diagonal positions, levels) depends on the structure only and is cached in the mpg_csr.
//
// Application (ilusv_jacobi): `steps` Jacobi sweeps for L (unit lower) and `steps` for U, each sweep ONE SpMV-shaped launch on
// the tuned packed kernel (sell.cu) with the update in its epilogue:
//     x' = x + (b - (I + L) x)            instead of copy + ilu_jacobi_mv + axpy            (kernels.hpp:236-240)
//     x' = x + D^-1 (b - U x)             instead of copy + ilu_jacobi_mv + gdmv            (kernels.hpp:244-248)
// on two CSR operators split out of the factors once: M_L = [L | 1] (unit diagonal stored last) and U (diagonal first).
#include <algorithm>
#include <vector>

#include "common.cuh"

using namespace mpg;

namespace mpg {
int scan_i32(mpg_ctx* ctx, int64_t n, const int* in, int* out);   // sell.cu: exclusive prefix sum, out[n] = total
int cast_copy(mpg_ctx*, int64_t, const double*, float*);
int cast_copy(mpg_ctx*, int64_t, const double*, double*);
int cast_copy(mpg_ctx*, int64_t, const float*, float*);
template <class T> int spmv(mpg_ctx*, const mpg_csr*, const T*, T, const T*, T, const T*, T*, float*, const T* rowscale, int part);
template <class T> int pack_create(mpg_ctx*, const mpg_csr*, const T*, mpg_packed**);
void pack_free(mpg_packed*);
template <class T> int spmv_packed(mpg_ctx*, const mpg_packed*, T, const T*, T, const T*, T*, float*, const T*, int, const HaloWait*, const T* xadd, const PushArgs* push = nullptr);
int axpy_host(mpg_ctx*, int64_t, float, const float*, float*);
int axpy_host(mpg_ctx*, int64_t, double, const double*, double*);
int gdmv_host(mpg_ctx*, int64_t, float, const float*, const float*, float, float*);
int gdmv_host(mpg_ctx*, int64_t, double, const double*, const double*, double, double*);
}  // namespace mpg

struct mpg_ilu_plan {
    int n = 0;
    int* diag_pos = nullptr;     // [n] position of the diagonal entry of every row in the CSR arrays
    int* level_rows = nullptr;   // [n] rows bucketed by level
    std::vector<int> level_ptr;  // [nlevels + 1] host
    int no_diag = 0;             // rows that store no diagonal entry (factorisation refused)
};

namespace {

__global__ void diag_pos_kernel(int n, const int* __restrict__ row_map, const int* __restrict__ inds, int* __restrict__ diag_pos, int* missing) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int j = row_map[i];
    const int je = row_map[i + 1];
    while (j < je && inds[j] < i) ++j;   // types.hpp:296-303 (bounded: a row without a diagonal is reported, not run past)
    diag_pos[i] = j;
    if (j >= je || inds[j] != i) atomicAdd(missing, 1);
}

// one relaxation sweep: level[i] = max over lower columns k of level[k] + 1 (0 without lower entries)
__global__ void level_sweep_kernel(int n, const int* __restrict__ row_map, const int* __restrict__ inds, const int* __restrict__ diag_pos, int* level,
                                   int* changed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lv = 0;
    for (int p = row_map[i], pe = diag_pos[i]; p < pe; ++p) lv = max(lv, *((volatile int*)(level + inds[p])) + 1);
    if (lv != level[i]) { level[i] = lv; *changed = 1; }
}
__global__ void level_hist_kernel(int n, const int* __restrict__ level, int* hist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(hist + level[i], 1);
}
__global__ void level_scatter_kernel(int n, const int* __restrict__ level, int* cursor, int* level_rows) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) level_rows[atomicAdd(cursor + level[i], 1)] = i;
}
__global__ void max_kernel(int n, const int* __restrict__ v, int* out) {
    int m = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, v[i]);
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// rows [first, first + count) of level_rows; warp per row
__global__ void __launch_bounds__(256) ilu0_level_kernel(int first, int count, const int* __restrict__ level_rows, const int* __restrict__ row_map,
                                                          const int* __restrict__ inds, const int* __restrict__ diag_pos, double* vals,
                                                          const double* __restrict__ amax, double eps_type) {
    const int w = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (w >= count) return;
    const int i = level_rows[first + w];
    const int re = row_map[i + 1], dp = diag_pos[i];
    for (int kk = row_map[i]; kk < dp; ++kk) {
        const int k = inds[kk];
        const int dk = diag_pos[k], ke = row_map[k + 1];
        const double factor = __ddiv_rn(__ldcg(vals + kk), __ldcg(vals + dk));   // kernels_mkl.cpp:457
        __syncwarp();                       // every lane has read a_ik before lane 0 overwrites it
        if (lane == 0) __stcg(vals + kk, factor);
        for (int p = dk + 1 + lane; p < ke; p += 32) {   // entries of row k right of its pivot
            const int c = inds[p];
            int lo = kk + 1, hi = re;       // binary search for column c among the remaining entries of row i (ascending columns)
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (inds[mid] < c) lo = mid + 1; else hi = mid;
            }
            if (lo < re && inds[lo] == c) __stcg(vals + lo, fma(-factor, __ldcg(vals + p), __ldcg(vals + lo)));   // :466-467
        }
        __syncwarp();                       // the updates of this k are visible to the whole warp before the next pivot is read
    }
    if (lane == 0 && i >= 1) {              // pivot boost, kernels_mkl.cpp:474-483 (the loop starts at row 1)
        const double alpha = *amax * eps_type;
        double d = __ldcg(vals + dp);
        if (d >= 0) { if (d < alpha) d = alpha; }
        else if (d > -alpha) d = -alpha;
        __stcg(vals + dp, d);
    }
}

__global__ void rowabs_max_f64_kernel(int nrows, const int* __restrict__ row_map, const double* __restrict__ vals, unsigned long long* out_bits) {
    double m = 0;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
        double s = 0;
        for (int p = row_map[r]; p < row_map[r + 1]; ++p) s += fabs(vals[p]);   // sequential, like the reference lambda (kernels_cuda.cpp:748-758)
        m = fmax(m, s);
    }
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, (unsigned long long)__double_as_longlong(m));   // non-negative doubles order like their bit patterns
}

// ---- split of the factors into M_L = [L | 1] and U --------------------------------------------------------------------------
__global__ void split_count_kernel(int n, const int* __restrict__ row_map, const int* __restrict__ diag_pos, int* nl, int* nu) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    nl[i] = diag_pos[i] - row_map[i] + 1;
    nu[i] = row_map[i + 1] - diag_pos[i];
}
template <class T>
__global__ void split_fill_kernel(int n, const int* __restrict__ row_map, const int* __restrict__ inds, const T* __restrict__ vals,
                                  const int* __restrict__ diag_pos, const int* __restrict__ rmL, const int* __restrict__ rmU, int* indL, T* valL, int* indU,
                                  T* valU, T* diag) {
    const int w = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (w >= n) return;
    const int rs = row_map[w], dp = diag_pos[w], re = row_map[w + 1];
    const int oL = rmL[w], oU = rmU[w];
    for (int p = rs + lane; p < dp; p += 32) { indL[oL + p - rs] = inds[p]; valL[oL + p - rs] = vals[p]; }
    for (int p = dp + lane; p < re; p += 32) { indU[oU + p - dp] = inds[p]; valU[oU + p - dp] = vals[p]; }
    if (lane == 0) {
        indL[oL + dp - rs] = w;
        valL[oL + dp - rs] = T(1);
        diag[w] = T(1) / vals[dp];   // types.hpp:301
    }
}

}  // namespace

namespace mpg {
void ilu_plan_free(mpg_ilu_plan* p) {
    if (!p) return;
    pool_free(p->diag_pos);
    pool_free(p->level_rows);
    delete p;
}

int ilu_plan_get(mpg_ctx* ctx, const mpg_csr* A, const mpg_ilu_plan** out) {
    mpg_csr* Am = const_cast<mpg_csr*>(A);
    if (Am->ilu) { *out = Am->ilu; return MPG_OK; }
    const int n = A->nrows;
    mpg_ilu_plan* p = new mpg_ilu_plan();
    struct Guard { mpg_ilu_plan* p; ~Guard() { if (p) ilu_plan_free(p); } } guard{p};
    p->n = n;
    int *level = nullptr, *flags = nullptr, *hist = nullptr;
    MPG_CUDA(ctx, pool_alloc(ctx, &p->diag_pos, sizeof(int) * (size_t)std::max(n, 1)));
    MPG_CUDA(ctx, pool_alloc(ctx, &p->level_rows, sizeof(int) * (size_t)std::max(n, 1)));
    MPG_CUDA(ctx, pool_alloc(ctx, &level, sizeof(int) * (size_t)std::max(n, 1)));
    MPG_CUDA(ctx, pool_alloc(ctx, &flags, sizeof(int) * 4));
    struct Tmp { int* a; int* b; int** c; ~Tmp() { pool_free(a); pool_free(b); if (*c) pool_free(*c); } } tmp{level, flags, &hist};
    MPG_CUDA(ctx, cudaMemsetAsync(level, 0, sizeof(int) * (size_t)std::max(n, 1), ctx->stream));
    MPG_CUDA(ctx, cudaMemsetAsync(flags, 0, sizeof(int) * 4, ctx->stream));
    const int grid = (int)cdiv(std::max(n, 1), 256);
    diag_pos_kernel<<<grid, 256, 0, ctx->stream>>>(n, A->row_map, A->inds, p->diag_pos, flags);
    MPG_CHECK_LAUNCH(ctx);
    MPG_CUDA(ctx, cudaMemcpyAsync(&p->no_diag, flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (p->no_diag > 0) return fail(ctx, MPG_ERR_ARG, "ilu0: " + std::to_string(p->no_diag) + " row(s) store no diagonal entry");
    // levels: relaxation sweeps, 4 per host check, until a whole batch changes nothing
    for (int64_t it = 0;; ++it) {
        if (it > (int64_t)n + 8) return fail(ctx, MPG_ERR_STATE, "ilu0: level relaxation did not converge");
        MPG_CUDA(ctx, cudaMemsetAsync(flags + 1, 0, sizeof(int), ctx->stream));
        for (int r = 0; r < 4; ++r) level_sweep_kernel<<<grid, 256, 0, ctx->stream>>>(n, A->row_map, A->inds, p->diag_pos, level, flags + 1);
        ctx->launches += 4;
        int changed = 0;
        MPG_CUDA(ctx, cudaMemcpyAsync(&changed, flags + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (!changed) break;
    }
    max_kernel<<<std::min(grid, 1024), 256, 0, ctx->stream>>>(n, level, flags + 2);
    MPG_CHECK_LAUNCH(ctx);
    int maxlev = 0;
    MPG_CUDA(ctx, cudaMemcpyAsync(&maxlev, flags + 2, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int nlev = maxlev + 1;
    MPG_CUDA(ctx, pool_alloc(ctx, &hist, sizeof(int) * (size_t)(nlev + 1)));
    MPG_CUDA(ctx, cudaMemsetAsync(hist, 0, sizeof(int) * (size_t)(nlev + 1), ctx->stream));
    level_hist_kernel<<<grid, 256, 0, ctx->stream>>>(n, level, hist);
    MPG_CHECK_LAUNCH(ctx);
    std::vector<int> h((size_t)nlev);
    MPG_CUDA(ctx, cudaMemcpyAsync(h.data(), hist, sizeof(int) * (size_t)nlev, cudaMemcpyDeviceToHost, ctx->stream));
    MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    p->level_ptr.assign((size_t)nlev + 1, 0);
    for (int l = 0; l < nlev; ++l) p->level_ptr[(size_t)l + 1] = p->level_ptr[(size_t)l] + h[(size_t)l];
    MPG_CUDA(ctx, cudaMemcpyAsync(hist, p->level_ptr.data(), sizeof(int) * (size_t)nlev, cudaMemcpyHostToDevice, ctx->stream));   // cursors
    level_scatter_kernel<<<grid, 256, 0, ctx->stream>>>(n, level, hist, p->level_rows);
    MPG_CHECK_LAUNCH(ctx);
    MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    guard.p = nullptr;
    Am->ilu = p;
    *out = p;
    return MPG_OK;
}
}  // namespace mpg

// ilu0<Type,Device>(SparseMatrix<double>) (kernels.hpp:163-164): factors of the fp64 matrix, stored on A's structure.
// eps_is_float selects numeric_limits<Type>::epsilon() for the boost threshold (Type = float / double).
extern "C" int mpg_ilu0_f64(mpg_ctx* ctx, const mpg_csr* A, const double* vals_in, int eps_is_float, double* vals_out) {
    MPG_REQUIRE(ctx, A && vals_in && vals_out, "ilu0: null argument");
    MPG_REQUIRE(ctx, A->nrows == A->ncols, "ilu0: the matrix must be square (a partitioned slab cannot be factored)");
    if (A->nrows == 0) return MPG_OK;
    const mpg_ilu_plan* p = nullptr;
    MPG_TRY(mpg::ilu_plan_get(ctx, A, &p));
    MPG_TRY(mpg::cast_copy(ctx, A->nnz, vals_in, vals_out));
    unsigned long long* amax = reinterpret_cast<unsigned long long*>(ctx->dscal + 24);
    MPG_CUDA(ctx, cudaMemsetAsync(amax, 0, sizeof(unsigned long long), ctx->stream));
    rowabs_max_f64_kernel<<<std::min((int)cdiv(A->nrows, 256), ctx->num_sms * 8), 256, 0, ctx->stream>>>(A->nrows, A->row_map, vals_in, amax);
    MPG_CHECK_LAUNCH(ctx);
    const double eps = eps_is_float ? 1.1920928955078125e-07 : 2.220446049250313e-16;
    const int nlev = (int)p->level_ptr.size() - 1;
    for (int l = 0; l < nlev; ++l) {
        const int first = p->level_ptr[(size_t)l], count = p->level_ptr[(size_t)l + 1] - first;
        if (count == 0) continue;
        ilu0_level_kernel<<<(int)cdiv((int64_t)count * 32, 256), 256, 0, ctx->stream>>>(first, count, p->level_rows, A->row_map, A->inds, p->diag_pos, vals_out,
                                                                                       reinterpret_cast<const double*>(amax), eps);
        MPG_CHECK_LAUNCH(ctx);
    }
    return MPG_OK;
}
extern "C" int mpg_ilu0_levels(mpg_ctx* ctx, const mpg_csr* A, int* nlevels) {
    MPG_REQUIRE(ctx, A && nlevels, "ilu0_levels: null argument");
    const mpg_ilu_plan* p = nullptr;
    MPG_TRY(mpg::ilu_plan_get(ctx, A, &p));
    *nlevels = (int)p->level_ptr.size() - 1;
    return MPG_OK;
}

// ---- ILU_Jacobi ---------------------------------------------------------------------------------------------------------
struct mpg_ilu_jacobi {
    int n = 0, steps = 1, tsize = 4, device = 0;
    int *rmL = nullptr, *indL = nullptr, *rmU = nullptr, *indU = nullptr;
    void *valL = nullptr, *valU = nullptr, *diag = nullptr, *temp1 = nullptr, *temp2 = nullptr, *vals = nullptr;
    mpg_csr *L = nullptr, *U = nullptr;
    mpg_packed *PL = nullptr, *PU = nullptr;
};

extern "C" int mpg_ilu_jacobi_destroy(mpg_ilu_jacobi* M) {
    if (!M) return MPG_OK;
    cudaSetDevice(M->device);
    mpg::pack_free(M->PL); mpg::pack_free(M->PU);
    mpg_csr_destroy(M->L); mpg_csr_destroy(M->U);
    pool_free(M->rmL); pool_free(M->indL); pool_free(M->rmU); pool_free(M->indU);
    pool_free(M->valL); pool_free(M->valU); pool_free(M->diag); pool_free(M->temp1); pool_free(M->temp2); pool_free(M->vals);
    delete M;
    return MPG_OK;
}

namespace {
template <class T>
int ilu_jacobi_create(mpg_ctx* ctx, const mpg_csr* A, const double* ilu_vals64, int steps, mpg_ilu_jacobi** out) {
    *out = nullptr;
    MPG_REQUIRE(ctx, A && ilu_vals64 && steps >= 0, "ilu_jacobi_create: bad argument");
    MPG_REQUIRE(ctx, A->nrows == A->ncols, "ilu_jacobi_create: the matrix must be square");
    const int n = A->nrows;
    const mpg_ilu_plan* p = nullptr;
    MPG_TRY(mpg::ilu_plan_get(ctx, A, &p));
    mpg_ilu_jacobi* M = new mpg_ilu_jacobi();
    struct Guard { mpg_ilu_jacobi* m; ~Guard() { if (m) mpg_ilu_jacobi_destroy(m); } } guard{M};
    M->n = n; M->steps = steps; M->tsize = (int)sizeof(T); M->device = ctx->device;
    const size_t nb = sizeof(int) * (size_t)(n + 1);
    int *nl = nullptr, *nu = nullptr;
    MPG_CUDA(ctx, pool_alloc(ctx, &nl, nb)); MPG_CUDA(ctx, pool_alloc(ctx, &nu, nb));
    struct Tmp { int* a; int* b; ~Tmp() { pool_free(a); pool_free(b); } } tmp{nl, nu};
    MPG_CUDA(ctx, pool_alloc(ctx, &M->rmL, nb)); MPG_CUDA(ctx, pool_alloc(ctx, &M->rmU, nb));
    MPG_CUDA(ctx, pool_alloc(ctx, &M->vals, sizeof(T) * (size_t)std::max<int64_t>(A->nnz, 1)));
    MPG_CUDA(ctx, pool_alloc(ctx, &M->diag, sizeof(T) * (size_t)std::max(n, 1)));
    MPG_CUDA(ctx, pool_alloc(ctx, &M->temp1, sizeof(T) * (size_t)std::max(n, 1)));
    MPG_CUDA(ctx, pool_alloc(ctx, &M->temp2, sizeof(T) * (size_t)std::max(n, 1)));
    MPG_TRY(mpg::cast_copy(ctx, A->nnz, ilu_vals64, static_cast<T*>(M->vals)));   // type_convert, kernels_cuda.cpp:697-712
    const int grid = (int)cdiv(std::max(n, 1), 256);
    split_count_kernel<<<grid, 256, 0, ctx->stream>>>(n, A->row_map, p->diag_pos, nl, nu);
    MPG_CHECK_LAUNCH(ctx);
    MPG_TRY(mpg::scan_i32(ctx, n, nl, M->rmL));
    MPG_TRY(mpg::scan_i32(ctx, n, nu, M->rmU));
    int nnzL = 0, nnzU = 0;
    MPG_CUDA(ctx, cudaMemcpyAsync(&nnzL, M->rmL + n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MPG_CUDA(ctx, cudaMemcpyAsync(&nnzU, M->rmU + n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    MPG_CUDA(ctx, pool_alloc(ctx, &M->indL, sizeof(int) * (size_t)std::max(nnzL, 1)));
    MPG_CUDA(ctx, pool_alloc(ctx, &M->indU, sizeof(int) * (size_t)std::max(nnzU, 1)));
    MPG_CUDA(ctx, pool_alloc(ctx, &M->valL, sizeof(T) * (size_t)std::max(nnzL, 1)));
    MPG_CUDA(ctx, pool_alloc(ctx, &M->valU, sizeof(T) * (size_t)std::max(nnzU, 1)));
    split_fill_kernel<T><<<(int)cdiv((int64_t)std::max(n, 1) * 32, 256), 256, 0, ctx->stream>>>(n, A->row_map, A->inds, static_cast<const T*>(M->vals), p->diag_pos, M->rmL,
                                                                                             M->rmU, M->indL, static_cast<T*>(M->valL), M->indU,
                                                                                             static_cast<T*>(M->valU), static_cast<T*>(M->diag));
    MPG_CHECK_LAUNCH(ctx);
    MPG_TRY(mpg_csr_create(ctx, n, n, nnzL, M->rmL, M->indL, &M->L));
    MPG_TRY(mpg_csr_create(ctx, n, n, nnzU, M->rmU, M->indU, &M->U));
    if (ctx->tune.spmv_packed) {
        MPG_TRY(mpg::pack_create<T>(ctx, M->L, static_cast<const T*>(M->valL), &M->PL));
        MPG_TRY(mpg::pack_create<T>(ctx, M->U, static_cast<const T*>(M->valU), &M->PU));
    }
    guard.m = nullptr;
    *out = M;
    return MPG_OK;
}

// one Jacobi sweep: xnew = xold + [d .*] (b - Op xold)
template <class T>
int sweep(mpg_ctx* ctx, const mpg_ilu_jacobi* M, bool lower, const T* b, const T* xold, T* xnew) {
    const mpg_packed* P = lower ? M->PL : M->PU;
    const T* d = lower ? nullptr : static_cast<const T*>(M->diag);
    if (P) return mpg::spmv_packed<T>(ctx, P, T(-1), xold, T(1), b, xnew, nullptr, d, SPMV_ALL, nullptr, xold);
    // structures that do not pack: CSR kernel + the reference's separate update pass
    const mpg_csr* Op = lower ? M->L : M->U;
    T* t = xnew;   // temp = b - Op xold
    MPG_TRY((mpg::spmv<T>(ctx, Op, static_cast<const T*>(lower ? M->valL : M->valU), T(-1), xold, T(1), b, t, nullptr, nullptr, SPMV_ALL)));
    if (lower) {   // xnew = xold + temp: temp already sits in xnew
        return mpg::axpy_host(ctx, M->n, T(1), xold, xnew);
    }
    // xnew = 1 * xold + (1 * d) * temp needs temp and xold apart: scale in place, then add
    MPG_TRY(mpg::gdmv_host(ctx, M->n, T(1), d, t, T(0), t));
    return mpg::axpy_host(ctx, M->n, T(1), xold, xnew);
}

// ilusv_jacobi, kernels.hpp:227-248
template <class T>
int ilu_jacobi_apply(mpg_ctx* ctx, mpg_ilu_jacobi* M, T* x) {
    MPG_REQUIRE(ctx, M && x && M->tsize == (int)sizeof(T), "ilu_jacobi_apply: null argument or wrong precision");
    if (M->n == 0 || M->steps == 0) return MPG_OK;
    T* b = static_cast<T*>(M->temp1);
    T* cur = x;
    T* oth = static_cast<T*>(M->temp2);
    MPG_TRY(mpg::cast_copy(ctx, M->n, (const T*)cur, b));
    for (int s = 0; s < M->steps; ++s) { MPG_TRY(sweep<T>(ctx, M, true, b, cur, oth)); std::swap(cur, oth); }
    MPG_TRY(mpg::cast_copy(ctx, M->n, (const T*)cur, b));
    for (int s = 0; s < M->steps; ++s) { MPG_TRY(sweep<T>(ctx, M, false, b, cur, oth)); std::swap(cur, oth); }
    // 2 * steps swaps: cur == x again
    return MPG_OK;
}

// ilu_jacobi_mv, kernels.hpp:172-216 (the upper form ignores alpha / beta exactly like the reference)
template <class T>
int ilu_jacobi_mv(mpg_ctx* ctx, const mpg_ilu_jacobi* M, int lower, T alpha, const T* x, T beta, T* y) {
    MPG_REQUIRE(ctx, M && x && y && M->tsize == (int)sizeof(T), "ilu_jacobi_mv: null argument or wrong precision");
    if (!lower) { alpha = T(-1); beta = T(1); }
    const mpg_packed* P = lower ? M->PL : M->PU;
    if (P) return mpg::spmv_packed<T>(ctx, P, alpha, x, beta, y, y, nullptr, nullptr, SPMV_ALL, nullptr, nullptr);
    return mpg::spmv<T>(ctx, lower ? M->L : M->U, static_cast<const T*>(lower ? M->valL : M->valU), alpha, x, beta, y, y, nullptr, nullptr, SPMV_ALL);
}
}  // namespace

namespace mpg {
template <class T> int ilu_jacobi_apply_t(mpg_ctx* ctx, mpg_ilu_jacobi* M, T* x) { return ilu_jacobi_apply<T>(ctx, M, x); }
template int ilu_jacobi_apply_t<float>(mpg_ctx*, mpg_ilu_jacobi*, float*);
template int ilu_jacobi_apply_t<double>(mpg_ctx*, mpg_ilu_jacobi*, double*);
}  // namespace mpg

extern "C" int mpg_ilu_jacobi_create_f32(mpg_ctx* ctx, const mpg_csr* A, const double* ilu_vals, int steps, mpg_ilu_jacobi** out) {
    MPG_REQUIRE(ctx, out != nullptr, "ilu_jacobi_create: null out");
    return ilu_jacobi_create<float>(ctx, A, ilu_vals, steps, out);
}
extern "C" int mpg_ilu_jacobi_create_f64(mpg_ctx* ctx, const mpg_csr* A, const double* ilu_vals, int steps, mpg_ilu_jacobi** out) {
    MPG_REQUIRE(ctx, out != nullptr, "ilu_jacobi_create: null out");
    return ilu_jacobi_create<double>(ctx, A, ilu_vals, steps, out);
}
extern "C" int mpg_ilu_jacobi_apply_f32(mpg_ctx* ctx, mpg_ilu_jacobi* M, float* x) { return ilu_jacobi_apply<float>(ctx, M, x); }
extern "C" int mpg_ilu_jacobi_apply_f64(mpg_ctx* ctx, mpg_ilu_jacobi* M, double* x) { return ilu_jacobi_apply<double>(ctx, M, x); }
extern "C" int mpg_ilu_jacobi_mv_f32(mpg_ctx* ctx, const mpg_ilu_jacobi* M, int lower, float alpha, const float* x, float beta, float* y) {
    return ilu_jacobi_mv<float>(ctx, M, lower, alpha, x, beta, y);
}
extern "C" int mpg_ilu_jacobi_mv_f64(mpg_ctx* ctx, const mpg_ilu_jacobi* M, int lower, double alpha, const double* x, double beta, double* y) {
    return ilu_jacobi_mv<double>(ctx, M, lower, alpha, x, beta, y);
}
