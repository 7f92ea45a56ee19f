// spmv.cu — CSR SpMV y = alpha*A*x + beta*y in fp32 / fp64, and the fused fp64 outer residual + cast.
// Reference surface: kernels.hpp:159-165; reference CUDA backend: cusparse?csrmv, kernels_cuda.cpp:576-614.
//
// Design (B200): the matrix stream (4 B index + s B value per nonzero) is ~95 % of the compulsory traffic,
// so the kernel is organised around streaming it perfectly, not around rows:
//   * the nonzero range is cut into fixed tiles of TILE = 2048 nonzeros (nnz-split, merge-path style);
//     tile starts are multiples of TILE, so every index/value load is a 16-byte aligned, fully coalesced
//     streaming load (ld.global.nc.L1::no_allocate.v4) regardless of the row structure -> regular
//     stencils and power-law rows run through the same code with perfect load balance;
//   * products val*x[col] go to shared memory; x is gathered through L1/L2 (stencil neighbours hit);
//   * rows are then reduced out of shared memory: thread-per-row when the tile holds many rows,
//     warp-per-row when it holds few long ones;
//   * rows that straddle a tile boundary leave partial sums in carry_in / carry_out; a tiny fix-up kernel
//     (one thread per tile) adds them in tile order, so results are deterministic (no float atomics).
// This is the operator-surface SpMV on raw CSR arrays (any structure, values passed per call) and the fp64 residual
// kernel; the solver's inner iterations and SparseMatrix<T,B200> multiply with the packed copy of sell.cu when the
// structure packs (the x gathers of this kernel - one per nonzero - make it L1TEX-tag bound at ~75 % of the roofline).
// The plan (tile -> first row) is built once per matrix structure on the device (mpg_csr_create), the
// analogue of the reference's create_cuda_handles (types_cuda.hpp:53-60); fp32 and fp64 value arrays
// share it exactly as SparseMatrix<float,Cuda> aliases row_map/inds (types_cuda.hpp:82-91).
#include <vector>

#include "common.cuh"

using namespace mpg;

namespace {

constexpr int SPMV_THREADS = 256;
constexpr int SPMV_NPT = 8;                          // nonzeros per thread
constexpr int SPMV_TILE = SPMV_THREADS * SPMV_NPT;   // 2048

// tile_row[t] = row containing nonzero t*TILE, i.e. the largest r with row_map[r] <= t*TILE (and
// row_map[r+1] > t*TILE: rows are never empty in canonical form, but empty rows are tolerated: we take
// the LAST row whose start is <= the offset, skipping empties).
__global__ void plan_kernel(int nrows, int64_t nnz, const int* __restrict__ row_map, int tile_nnz, int ntiles, int* tile_row) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    if (t == ntiles) { tile_row[t] = nrows; return; }
    const int64_t off = (int64_t)t * tile_nnz;
    // upper_bound(row_map, off) - 1
    int lo = 0, hi = nrows + 1;  // search in row_map[0..nrows]
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((int64_t)row_map[mid] <= off) lo = mid + 1; else hi = mid;
    }
    tile_row[t] = lo - 1;
}

template <class T> struct Carry { using type = T; };

// gdmv(1, d, v, 0, v) of kernels.hpp:143-145 with the same rounding sequence as the stand-alone kernel in blas1.cu
__device__ __forceinline__ float rowscale_apply(float d, float v) { return __fadd_rn(__fmul_rn(0.f, v), __fmul_rn(__fmul_rn(1.f, d), v)); }
__device__ __forceinline__ double rowscale_apply(double d, double v) { return __dadd_rn(__dmul_rn(0.0, v), __dmul_rn(__dmul_rn(1.0, d), v)); }

// Epilogue: y_out[r] = alpha*sum + beta*y_in[r]  (beta == 0: y_in never read), optional Jacobi row scaling
// (the preconditioner application M(w) = diag .* w of types.hpp:444-446 folded into the store), optional fp32 copy.
template <class T>
__device__ __forceinline__ void spmv_store(int r, T sum, T alpha, T beta, const T* y_in, T* y_out, float* out32, const T* rowscale) {
    T v = (beta == T(0)) ? alpha * sum : fma(alpha, sum, beta * y_in[r]);
    if (rowscale) v = rowscale_apply(__ldg(rowscale + r), v);
    if (y_out) y_out[r] = v;
    if (out32) out32[r] = (float)v;
}

template <class T>
__global__ void __launch_bounds__(SPMV_THREADS) spmv_tile_kernel(int nrows, int64_t nnz, const int* __restrict__ row_map,
                                                                  const int* __restrict__ inds, const T* __restrict__ vals,
                                                                  const int* __restrict__ tile_row, const T* __restrict__ x,
                                                                  T alpha, T beta, const T* y_in, T* y_out, float* out32,
                                                                  T* carry_in, T* carry_out, int vec_ok, const T* rowscale,
                                                                  const int* __restrict__ tile_list) {
    __shared__ T prod[SPMV_TILE];
    __shared__ int rm_s[SPMV_TILE + 2];

    const int t = tile_list ? __ldg(tile_list + blockIdx.x) : blockIdx.x;
    const int64_t base = (int64_t)t * SPMV_TILE;
    const int64_t end = min(nnz, base + SPMV_TILE);
    const int cnt = (int)(end - base);
    const int r_lo = tile_row[t];
    // last row touched by this tile: the row containing nonzero end-1
    int r_hi = tile_row[t + 1];
    if (r_hi >= nrows) r_hi = nrows - 1;
    else if ((int64_t)__ldg(row_map + r_hi) >= end) r_hi -= 1;  // next tile starts exactly at a row start
    // skip trailing rows that begin at or beyond `end` (possible only with empty rows)
    const int nr = r_hi - r_lo + 1;

    // ---- stage row_map[r_lo .. r_hi+1] ----
    for (int i = threadIdx.x; i <= nr; i += SPMV_THREADS) rm_s[i] = __ldg(row_map + r_lo + i);

    // ---- stream indices + values, gather x, form products ----
    constexpr int VPT = 16 / sizeof(T);          // values per 16-byte load (4 fp32 / 2 fp64)
    const int tid = threadIdx.x;
    if (cnt == SPMV_TILE && vec_ok) {
        const int4* ip = reinterpret_cast<const int4*>(inds + base);
        int4 c0 = ldg_stream(ip + tid);
        int4 c1 = ldg_stream(ip + SPMV_THREADS + tid);
        if (sizeof(T) == 4) {
            const float4* vp = reinterpret_cast<const float4*>(vals + base);
            const float4 v0 = ldg_stream(vp + tid);
            const float4 v1 = ldg_stream(vp + SPMV_THREADS + tid);
            const T x0 = __ldg(x + c0.x), x1 = __ldg(x + c0.y), x2 = __ldg(x + c0.z), x3 = __ldg(x + c0.w);
            const T x4 = __ldg(x + c1.x), x5 = __ldg(x + c1.y), x6 = __ldg(x + c1.z), x7 = __ldg(x + c1.w);
            float4 p0, p1;
            p0.x = v0.x * x0; p0.y = v0.y * x1; p0.z = v0.z * x2; p0.w = v0.w * x3;
            p1.x = v1.x * x4; p1.y = v1.y * x5; p1.z = v1.z * x6; p1.w = v1.w * x7;
            reinterpret_cast<float4*>(prod)[tid] = p0;
            reinterpret_cast<float4*>(prod)[SPMV_THREADS + tid] = p1;
        } else {
            const double2* vp = reinterpret_cast<const double2*>(vals + base);
            // element e = q*1024 + tid*4 + c for the index loads; values use 2-wide loads on the same e
            const double2 v00 = ldg_stream(vp + 2 * tid), v01 = ldg_stream(vp + 2 * tid + 1);
            const double2 v10 = ldg_stream(vp + 2 * (SPMV_THREADS + tid)), v11 = ldg_stream(vp + 2 * (SPMV_THREADS + tid) + 1);
            const T x0 = __ldg(x + c0.x), x1 = __ldg(x + c0.y), x2 = __ldg(x + c0.z), x3 = __ldg(x + c0.w);
            const T x4 = __ldg(x + c1.x), x5 = __ldg(x + c1.y), x6 = __ldg(x + c1.z), x7 = __ldg(x + c1.w);
            double2* pp = reinterpret_cast<double2*>(prod);
            pp[2 * tid] = make_double2(v00.x * x0, v00.y * x1);
            pp[2 * tid + 1] = make_double2(v01.x * x2, v01.y * x3);
            pp[2 * (SPMV_THREADS + tid)] = make_double2(v10.x * x4, v10.y * x5);
            pp[2 * (SPMV_THREADS + tid) + 1] = make_double2(v11.x * x6, v11.y * x7);
        }
    } else {
        for (int e = tid; e < cnt; e += SPMV_THREADS) prod[e] = ldg_stream(vals + base + e) * __ldg(x + ldg_stream(inds + base + e));
    }
    (void)VPT;
    __syncthreads();

    // ---- reduce rows out of shared memory ----
    const bool many_rows = nr > (SPMV_THREADS / 8);
    if (many_rows) {
        for (int lr = tid; lr < nr; lr += SPMV_THREADS) {
            const int64_t rs = rm_s[lr], re = rm_s[lr + 1];
            const int s = (int)(max(rs, base) - base), e = (int)(min(re, end) - base);
            T sum = T(0);
            for (int p = s; p < e; ++p) sum += prod[p];
            const bool head_cut = rs < base, tail_cut = re > end;
            if (!head_cut && !tail_cut) spmv_store<T>(r_lo + lr, sum, alpha, beta, y_in, y_out, out32, rowscale);
            else if (head_cut) carry_in[t] = sum;       // continues a row begun in an earlier tile
            else carry_out[t] = sum;                    // row begins here, finishes later
        }
    } else {
        const int warp = tid >> 5, lane = tid & 31;
        for (int lr = warp; lr < nr; lr += SPMV_THREADS / 32) {
            const int64_t rs = rm_s[lr], re = rm_s[lr + 1];
            const int s = (int)(max(rs, base) - base), e = (int)(min(re, end) - base);
            T sum = T(0);
            for (int p = s + lane; p < e; p += 32) sum += prod[p];
            sum = warp_sum(sum);
            if (lane == 0) {
                const bool head_cut = rs < base, tail_cut = re > end;
                if (!head_cut && !tail_cut) spmv_store<T>(r_lo + lr, sum, alpha, beta, y_in, y_out, out32, rowscale);
                else if (head_cut) carry_in[t] = sum;
                else carry_out[t] = sum;
            }
        }
    }
}

// One thread per tile: if a row BEGINS in tile t and is cut by its end, gather the pieces in tile order.
template <class T>
__global__ void spmv_fixup_kernel(int nrows, int64_t nnz, int ntiles, const int* __restrict__ row_map,
                                  const int* __restrict__ tile_row, T alpha, T beta, const T* y_in, T* y_out, float* out32,
                                  const T* carry_in, const T* carry_out, const T* rowscale) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntiles - 1) return;  // the last tile cannot be cut at its end
    const int64_t end = (int64_t)(t + 1) * SPMV_TILE;
    const int r = tile_row[t + 1];  // row containing the first nonzero of tile t+1
    if (r >= nrows) return;
    const int64_t rs = row_map[r], re = row_map[r + 1];
    if (rs >= end) return;                       // tile t+1 starts exactly at a row start: nothing is cut
    if (rs < end - SPMV_TILE) return;            // row began before tile t: an earlier thread owns it
    T sum = carry_out[t];
    for (int t2 = t + 1; t2 < ntiles; ++t2) {
        sum += carry_in[t2];
        if (re <= (int64_t)(t2 + 1) * SPMV_TILE) break;
    }
    spmv_store<T>(r, sum, alpha, beta, y_in, y_out, out32, rowscale);
}

// Non-canonical input (rows with no stored entry at all; LoadMatrix-canonical matrices always hold the diagonal,
// LoadMatrix.hpp:98-100): the tile plan assumes every row owns at least one nonzero, so such matrices take this
// plain warp-per-row kernel instead.
template <class T>
__global__ void __launch_bounds__(256) spmv_rows_kernel(int nrows, const int* __restrict__ row_map, const int* __restrict__ inds,
                                                         const T* __restrict__ vals, const T* __restrict__ x, T alpha, T beta,
                                                         const T* y_in, T* y_out, float* out32, const T* rowscale) {
    const int lane = threadIdx.x & 31;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < nrows; r += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const int rs = __ldg(row_map + r), re = __ldg(row_map + r + 1);
        T sum = T(0);
        for (int p = rs + lane; p < re; p += 32) sum += ldg_stream(vals + p) * __ldg(x + ldg_stream(inds + p));
        sum = warp_sum(sum);
        if (lane == 0) spmv_store<T>((int)r, sum, alpha, beta, y_in, y_out, out32, rowscale);
    }
}

template <class T>
int launch_spmv(mpg_ctx* ctx, const mpg_csr* A, const T* vals, T alpha, const T* x, T beta, const T* y_in, T* y_out, float* out32,
                const T* rowscale = nullptr, int part = SPMV_ALL) {
    if (A->nrows == 0) return MPG_OK;
    // part: the solver splits a partitioned SpMV into the tiles without halo columns (SPMV_INTERIOR, may run before the
    // halo has arrived) and the rest + the cut-row fix-up (SPMV_BOUNDARY); without a tile list the second call does it all
    if (part != SPMV_ALL && (!A->tile_list || A->has_empty_rows)) {
        if (part == SPMV_INTERIOR) return MPG_OK;
        part = SPMV_ALL;
    }
    const int t_first = part == SPMV_BOUNDARY ? A->n_interior_tiles : 0;
    const int t_count = part == SPMV_INTERIOR ? A->n_interior_tiles : A->ntiles - t_first;
    const int* tile_list = part == SPMV_ALL ? nullptr : A->tile_list + t_first;
    const double share = A->ntiles > 0 ? (double)t_count / A->ntiles : 1.0;
    T* carry_in = reinterpret_cast<T*>(A->carry);
    T* carry_out = carry_in + A->ntiles;
    // algorithmic bytes, SURVEY.md §8d: nnz*(s+4) + 4(n+1) + n*s (x) + n*s (y)  [+ n*s for y_in when beta != 0, + 4n for the fp32 copy]
    const double n_ = A->nrows, s_ = sizeof(T);
    const double bytes = (double)A->nnz * (s_ + 4) + 4 * (n_ + 1) + n_ * s_ + (y_out ? n_ * s_ : 0) + (beta != T(0) ? n_ * s_ : 0) + (out32 ? 4 * n_ : 0) +
                         (rowscale ? n_ * s_ : 0);
    if (part == SPMV_INTERIOR && t_count == 0) return MPG_OK;
    ProfScope prof(ctx, sizeof(T) == 4 ? MPG_PROF_SPMV_F32 : MPG_PROF_SPMV_F64, bytes * share);
    if (A->has_empty_rows) {
        const int grid = (int)std::min<int64_t>(cdiv((int64_t)A->nrows * 32, 256), (int64_t)ctx->num_sms * 32);
        spmv_rows_kernel<T><<<grid, 256, 0, ctx->stream>>>(A->nrows, A->row_map, A->inds, vals, x, alpha, beta, y_in, y_out, out32, rowscale);
        MPG_CHECK_LAUNCH(ctx);
    } else if (A->ntiles > 0) {
        // 16-byte vector loads need aligned index / value arrays; sub-views fall back to scalar streaming loads
        const int vec_ok = ((reinterpret_cast<uintptr_t>(A->inds) | reinterpret_cast<uintptr_t>(vals)) & 15) == 0;
        // (two "gather x per row from staged indices" variants were measured and dropped: 0.93 / 1.10 ms against 0.75 ms
        //  on cd27:256 - profiles/r01_tune_spmv_variants.txt)
        if (t_count > 0) {
            spmv_tile_kernel<T><<<t_count, SPMV_THREADS, 0, ctx->stream>>>(A->nrows, A->nnz, A->row_map, A->inds, vals, A->tile_row, x, alpha, beta,
                                                                          y_in, y_out, out32, carry_in, carry_out, vec_ok, rowscale, tile_list);
            MPG_CHECK_LAUNCH(ctx);
        }
        if (A->ntiles > 1 && part != SPMV_INTERIOR) {
            spmv_fixup_kernel<T><<<(int)cdiv(A->ntiles - 1, 256), 256, 0, ctx->stream>>>(A->nrows, A->nnz, A->ntiles, A->row_map, A->tile_row,
                                                                                      alpha, beta, y_in, y_out, out32, carry_in, carry_out, rowscale);
            MPG_CHECK_LAUNCH(ctx);
        }
    }
    return MPG_OK;
}

// one warp per tile: does the tile reference a halo column (index >= nrows)?
__global__ void tile_halo_flag_kernel(int ntiles, int64_t nnz, int nrows, const int* __restrict__ inds, int* __restrict__ flag) {
    const int t = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (t >= ntiles) return;
    const int64_t base = (int64_t)t * SPMV_TILE, end = min(nnz, base + SPMV_TILE);
    int any = 0;
    for (int64_t p = base + lane; p < end; p += 32) any |= __ldg(inds + p) >= nrows;
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) flag[t] = any;
}

__global__ void count_empty_kernel(int nrows, const int* __restrict__ row_map, int* count) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nrows && row_map[r] == row_map[r + 1]) atomicAdd(count, 1);
}

}  // namespace

extern "C" int mpg_csr_create(mpg_ctx* ctx, int nrows, int ncols, int64_t nnz, const int* row_map, const int* inds, mpg_csr** out) {
    MPG_REQUIRE(ctx, out != nullptr, "csr_create: null out");
    MPG_REQUIRE(ctx, nrows >= 0 && ncols >= 0 && nnz >= 0, "csr_create: negative dims");
    MPG_REQUIRE(ctx, nnz < (int64_t)2147483647, "csr_create: nnz must fit int32 (types_cuda.hpp:66-70)");
    mpg_csr* A = new mpg_csr();
    struct Guard { mpg_csr* a; ~Guard() { if (a) mpg_csr_destroy(a); } } guard{A};   // error paths release what was built so far
    A->nrows = nrows; A->ncols = ncols; A->nnz = nnz; A->row_map = row_map; A->inds = inds;
    A->tile_nnz = SPMV_TILE;
    A->ntiles = (int)cdiv(nnz, SPMV_TILE);
    A->device = ctx->device;
    MPG_CUDA(ctx, pool_alloc(ctx, &A->tile_row, sizeof(int) * (size_t)(A->ntiles + 2)));
    MPG_CUDA(ctx, pool_alloc(ctx, &A->carry, sizeof(double) * 2 * (size_t)(A->ntiles + 1)));
    MPG_CUDA(ctx, cudaMemsetAsync(A->carry, 0, sizeof(double) * 2 * (size_t)(A->ntiles + 1), ctx->stream));
    plan_kernel<<<(int)cdiv(A->ntiles + 1, 256), 256, 0, ctx->stream>>>(nrows, nnz, row_map, SPMV_TILE, A->ntiles, A->tile_row);
    MPG_CHECK_LAUNCH(ctx);
    if (nrows > 0) {   // the scratch slot after the plan doubles as the empty-row counter
        int* cnt = A->tile_row + A->ntiles + 1;
        MPG_CUDA(ctx, cudaMemsetAsync(cnt, 0, sizeof(int), ctx->stream));
        count_empty_kernel<<<(int)cdiv(nrows, 256), 256, 0, ctx->stream>>>(nrows, row_map, cnt);
        MPG_CHECK_LAUNCH(ctx);
        int h = 0;
        MPG_CUDA(ctx, cudaMemcpyAsync(&h, cnt, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        A->has_empty_rows = h > 0;
    }
    if (ncols > nrows && A->ntiles > 0 && !A->has_empty_rows) {
        // local slab of a partitioned matrix: order the tiles [no halo column | some halo column]
        std::vector<int> flag((size_t)A->ntiles), list((size_t)A->ntiles);
        MPG_CUDA(ctx, pool_alloc(ctx, &A->tile_list, sizeof(int) * (size_t)A->ntiles));
        tile_halo_flag_kernel<<<(int)cdiv((int64_t)A->ntiles * 32, 256), 256, 0, ctx->stream>>>(A->ntiles, nnz, nrows, inds, A->tile_list);
        MPG_CHECK_LAUNCH(ctx);
        MPG_CUDA(ctx, cudaMemcpyAsync(flag.data(), A->tile_list, sizeof(int) * flag.size(), cudaMemcpyDeviceToHost, ctx->stream));
        MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        int ni = 0;
        for (int t = 0; t < A->ntiles; ++t) if (!flag[(size_t)t]) list[(size_t)ni++] = t;
        A->n_interior_tiles = ni;
        for (int t = 0; t < A->ntiles; ++t) if (flag[(size_t)t]) list[(size_t)ni++] = t;
        MPG_CUDA(ctx, cudaMemcpyAsync(A->tile_list, list.data(), sizeof(int) * list.size(), cudaMemcpyHostToDevice, ctx->stream));
        MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    guard.a = nullptr;
    *out = A;
    return MPG_OK;
}

extern "C" int mpg_csr_destroy(mpg_csr* A) {
    if (!A) return MPG_OK;
    cudaSetDevice(A->device);
    mpg::pool_free(A->tile_row);
    mpg::pool_free(A->carry);
    mpg::pool_free(A->tile_list);
    mpg::sell_plan_free(A->sell);
    mpg::ilu_plan_free(A->ilu);
    delete A;
    return MPG_OK;
}

namespace mpg {
template <class T>
int spmv(mpg_ctx* ctx, const mpg_csr* A, const T* vals, T alpha, const T* x, T beta, const T* y_in, T* y_out, float* out32, const T* rowscale,
         int part) {
    return launch_spmv<T>(ctx, A, vals, alpha, x, beta, y_in, y_out, out32, rowscale, part);
}
template int spmv<float>(mpg_ctx*, const mpg_csr*, const float*, float, const float*, float, const float*, float*, float*, const float*, int);
template int spmv<double>(mpg_ctx*, const mpg_csr*, const double*, double, const double*, double, const double*, double*, float*, const double*, int);
}  // namespace mpg

extern "C" int mpg_spmv_f32(mpg_ctx* ctx, const mpg_csr* A, const float* vals, float alpha, const float* x, float beta, float* y) {
    MPG_REQUIRE(ctx, A && (vals || A->nnz == 0) && x && y, "spmv: null argument");
    return launch_spmv<float>(ctx, A, vals, alpha, x, beta, y, y, nullptr);
}
extern "C" int mpg_spmv_f64(mpg_ctx* ctx, const mpg_csr* A, const double* vals, double alpha, const double* x, double beta, double* y) {
    MPG_REQUIRE(ctx, A && (vals || A->nnz == 0) && x && y, "spmv: null argument");
    return launch_spmv<double>(ctx, A, vals, alpha, x, beta, y, y, nullptr);
}
extern "C" int mpg_spmv_jacobi_f32(mpg_ctx* ctx, const mpg_csr* A, const float* vals, const float* diag, const float* x, float* y) {
    MPG_REQUIRE(ctx, A && vals && diag && x && y, "spmv_jacobi: null argument");
    return launch_spmv<float>(ctx, A, vals, 1.f, x, 0.f, y, y, nullptr, diag);
}
extern "C" int mpg_spmv_jacobi_f64(mpg_ctx* ctx, const mpg_csr* A, const double* vals, const double* diag, const double* x, double* y) {
    MPG_REQUIRE(ctx, A && vals && diag && x && y, "spmv_jacobi: null argument");
    return launch_spmv<double>(ctx, A, vals, 1.0, x, 0.0, y, y, nullptr, diag);
}
extern "C" int mpg_residual_f64_cast_f32(mpg_ctx* ctx, const mpg_csr* A, const double* vals, const double* b, const double* x,
                                         double* r64, float* w32) {
    MPG_REQUIRE(ctx, A && vals && b && x && w32, "residual: null argument");
    // r = b; r = -1*A*x + 1*r (gmres.cpp:173-174); w = (float) r (gmres.cpp:175)
    return launch_spmv<double>(ctx, A, vals, -1.0, x, 1.0, b, r64, w32);
}

// ---- Jacobi diagonal: types.hpp:395-430 ---------------------------------------------------------------
namespace {
template <class T>
__global__ void rowabs_max_kernel(int nrows, const int* __restrict__ row_map, const T* __restrict__ vals, double* partials,
                                  unsigned int* ticket, Epi epi) {
    T m = T(0);
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x) {
        T s = T(0);
        for (int p = row_map[r]; p < row_map[r + 1]; ++p) s += fabs(vals[p]);   // sequential, like the reference lambda
        m = max(m, s);
    }
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ T wm[8];
    if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = max(m, wm[w]);
        partials[blockIdx.x] = (double)m;
    }
    if (grid_last_block(ticket)) {
        __shared__ double gmax_s[1];
        if (threadIdx.x == 0) {
            double g = 0;
            for (unsigned b = 0; b < gridDim.x; ++b) g = max(g, __ldcg(partials + b));
            gmax_s[0] = g;
        }
        __syncthreads();
        finish_reduction<T>(epi, 1, gmax_s);
    }
}
template <class T>
__global__ void jacobi_diag_kernel(int nrows, const int* __restrict__ row_map, const int* __restrict__ inds, const T* __restrict__ vals,
                                   const T* amax, T* diag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nrows) return;
    const T alpha = *amax * (T)1.1920928955078125e-07f;  // numeric_limits<float>::epsilon(), types.hpp:416
    // the diagonal entry: LoadMatrix-canonical rows always hold it (LoadMatrix.hpp:62-66).  types.hpp:424-427 scans
    // for the first column >= i, which is the same entry on a sorted row; an equality scan also works on a
    // partitioned slab whose remote columns are renumbered past the local ones.
    // A row that stores no diagonal (raw CSR handed to mpg_csr_create is not checked) takes the small-pivot branch, i.e. v = 0.
    T v = T(0);
    for (int j = row_map[i], je = row_map[i + 1]; j < je; ++j)
        if (inds[j] == i) { v = vals[j]; break; }
    if (v >= 0) diag[i] = T(1) / ((v < alpha) ? alpha : v);
    else diag[i] = T(1) / ((v > -alpha) ? -alpha : v);
}
template <class T>
int jacobi_diag(mpg_ctx* ctx, const mpg_csr* A, const T* vals, T* diag) {
    if (A->nrows == 0) return MPG_OK;
    T* amax = reinterpret_cast<T*>(ctx->dscal + 8);
    const int grid = std::min<int>((int)cdiv(A->nrows, 256), ctx->num_sms * 8);
    const Epi epi = make_epi(ctx, EPI_MAX, amax, nullptr, 0.0, 0.0);
    rowabs_max_kernel<T><<<grid, 256, 0, ctx->stream>>>(A->nrows, A->row_map, vals, ctx->partials, ctx->ticket, epi);
    MPG_CHECK_LAUNCH(ctx);
    MPG_TRY(dist_finish_reduction(ctx, epi, 1, (int)sizeof(T)));
    jacobi_diag_kernel<T><<<(int)cdiv(A->nrows, 256), 256, 0, ctx->stream>>>(A->nrows, A->row_map, A->inds, vals, amax, diag);
    MPG_CHECK_LAUNCH(ctx);
    return MPG_OK;
}
}  // namespace
extern "C" int mpg_jacobi_diag_f32(mpg_ctx* ctx, const mpg_csr* A, const float* vals, float* diag) { return jacobi_diag<float>(ctx, A, vals, diag); }
extern "C" int mpg_jacobi_diag_f64(mpg_ctx* ctx, const mpg_csr* A, const double* vals, double* diag) { return jacobi_diag<double>(ctx, A, vals, diag); }
