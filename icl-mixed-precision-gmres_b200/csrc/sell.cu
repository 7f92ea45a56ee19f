// sell.cu — packed SpMV operator: the CSR matrix re-laid as 32-row slices (sliced ELLPACK, SELL-32).
// Reference surface: the same y = alpha*A*x + beta*y of kernels.hpp:159-165; the packed form plays the role of the
// reference's per-matrix library handles (create_cuda_handles, types_cuda.hpp:53-60: cusparse descriptors built once
// per SparseMatrix and re-used by every spmv call).
//
// Why: the CSR kernel (spmv.cu) streams the matrix perfectly but gathers x with one thread per NONZERO, so one warp
// gather touches ~10 different 128-byte lines of x (ncu: the kernel is L1TEX-tag bound at 73-77 % of the HBM
// roofline, profiles/r01_ncu_summary.md).  In the packed layout one warp owns 32 consecutive rows and lane = row:
//   * the p-th nonzeros of the 32 rows are adjacent in memory, G at a time per lane (G = 16 B / sizeof(T)), so
//     indices and values are still read with fully coalesced 16-byte streaming loads;
//   * a warp gather reads x[col_p(row)] for 32 consecutive rows - for stencil-like matrices 32 nearly consecutive
//     entries of x: 1-2 lines instead of ~10;
//   * every row is summed by its own lane in nonzero order: no shared memory, no cross-thread reduction, no
//     rows cut by tile boundaries, no fix-up kernel.
// Layout: slice s covers rows [32 s, 32 s + 32) and holds 32 * L_s elements, L_s = longest row of the slice.  The
// first floor(L_s / G) * G positions are stored in groups of G per lane: element (row r, position p) lives at
// slice_off[s] + (p / G) * 32 * G + (r % 32) * G + p % G; the remaining L_s % G positions are stored one per lane:
// slice_off[s] + floor(L_s / G) * 32 * G + (p % G) * 32 + r % 32.  (L_s itself is rounded up to a multiple of G when
// that costs at most 10 % padding.)  Shorter rows are padded by repeating their first column with value 0.  Matrices whose row lengths vary too much inside slices (padding > 25 %: the
// power-law case) are not packed - callers keep the CSR kernel.
// The packed INDICES depend only on the structure and are cached in the mpg_csr plan (one per G); the packed VALUES
// belong to an mpg_packed object and are refreshed with mpg_pack_update.
#include <vector>

#include "common.cuh"

using namespace mpg;

namespace {

constexpr int SLICE = 32;

template <class T> struct Grp;
template <> struct Grp<float> { static constexpr int G = 4; using IV = int4; using VV = float4; };
template <> struct Grp<double> { static constexpr int G = 2; using IV = int2; using VV = double2; };

__device__ __forceinline__ int2 ldg_stream2(const int2* p) {
    int2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ int4 ldg_idx(const int4* p) { return ldg_stream(p); }
__device__ __forceinline__ int2 ldg_idx(const int2* p) { return ldg_stream2(p); }

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }

// slice_len[s] = 32 * L_s (elements); L_s = longest row, rounded up to a whole group when that costs <= 10 % padding
// (one more 16-byte group is cheaper than up to G-1 single-element loads: cd27 27 -> 28; lap2d stays at 5 = 4 + 1)
__global__ void sell_len_kernel(int nrows, int nslices, const int* __restrict__ row_map, int G, int64_t* __restrict__ slice_len, int* has_rem) {
    const int s = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (s >= nslices) return;
    const int r = s * SLICE + lane;
    int len = (r < nrows) ? __ldg(row_map + r + 1) - __ldg(row_map + r) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    const int rem = len % G;
    if (rem > 0 && (G - rem) * 10 <= len) len += G - rem;
    if (lane == 0) {
        slice_len[s] = (int64_t)len * SLICE;
        if (len % G) *has_rem = 1;   // benign race: every writer stores 1
    }
}

// exclusive prefix sum of n slice lengths, out[n] = total.  One 1024-thread block: per-thread chunk sums, Hillis-Steele
// over the 1024 chunk sums, per-thread chunk write-out.  Plan-time only (4 MB at 16.7 M rows).
__global__ void __launch_bounds__(1024) sell_scan_kernel(int n, const int64_t* __restrict__ in, int64_t* __restrict__ out) {
    __shared__ int64_t part[1024];
    const int t = threadIdx.x;
    const int chunk = (n + 1023) / 1024;
    const int lo = min(t * chunk, n), hi = min(lo + chunk, n);
    int64_t s = 0;
    for (int i = lo; i < hi; ++i) s += in[i];
    part[t] = s;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
        const int64_t v = t >= off ? part[t - off] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    int64_t run = t ? part[t - 1] : 0;
    for (int i = lo; i < hi; ++i) {
        out[i] = run;
        run += in[i];
    }
    if (t == 1023) out[n] = part[1023];
}

// packed indices (structure); warp = slice, lane = row.  halo_flag[s] = 1 if the slice references a column >= nrows
template <int G>
__global__ void sell_fill_inds_kernel(int nrows, int nslices, const int* __restrict__ row_map, const int* __restrict__ inds,
                                      const int64_t* __restrict__ slice_off, int* __restrict__ sinds, int* __restrict__ halo_flag) {
    const int s = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (s >= nslices) return;
    const int r = s * SLICE + lane;
    const int64_t off = slice_off[s];
    const int L = (int)((slice_off[s + 1] - off) / SLICE);
    const int rs = (r < nrows) ? __ldg(row_map + r) : 0;
    const int len = (r < nrows) ? __ldg(row_map + r + 1) - rs : 0;
    const int pad = len > 0 ? __ldg(inds + rs) : 0;
    int any_halo = 0;
    const int ng = L / G;
    for (int g = 0; g < ng; ++g) {
        int c[G];
#pragma unroll
        for (int q = 0; q < G; ++q) {
            const int p = g * G + q;
            c[q] = p < len ? __ldg(inds + rs + p) : pad;
            any_halo |= c[q] >= nrows;
        }
        int* dst = sinds + off + (int64_t)g * SLICE * G + lane * G;
#pragma unroll
        for (int q = 0; q < G; ++q) dst[q] = c[q];
    }
    for (int p = ng * G; p < L; ++p) {
        const int c = p < len ? __ldg(inds + rs + p) : pad;
        any_halo |= c >= nrows;
        sinds[off + (int64_t)ng * SLICE * G + (int64_t)(p - ng * G) * SLICE + lane] = c;
    }
    any_halo = __any_sync(0xffffffffu, any_halo);
    if (lane == 0 && halo_flag) halo_flag[s] = any_halo;
}

template <class T>
__global__ void sell_fill_vals_kernel(int nrows, int nslices, const int* __restrict__ row_map, const T* __restrict__ vals,
                                      const int64_t* __restrict__ slice_off, T* __restrict__ svals) {
    constexpr int G = Grp<T>::G;
    const int s = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (s >= nslices) return;
    const int r = s * SLICE + lane;
    const int64_t off = slice_off[s];
    const int L = (int)((slice_off[s + 1] - off) / SLICE);
    const int rs = (r < nrows) ? __ldg(row_map + r) : 0;
    const int len = (r < nrows) ? __ldg(row_map + r + 1) - rs : 0;
    const int ng = L / G;
    for (int g = 0; g < ng; ++g) {
        typename Grp<T>::VV o;
        T* po = reinterpret_cast<T*>(&o);
#pragma unroll
        for (int q = 0; q < G; ++q) {
            const int p = g * G + q;
            po[q] = p < len ? __ldg(vals + rs + p) : T(0);
        }
        *reinterpret_cast<typename Grp<T>::VV*>(svals + off + (int64_t)g * SLICE * G + lane * G) = o;
    }
    for (int p = ng * G; p < L; ++p) svals[off + (int64_t)ng * SLICE * G + (int64_t)(p - ng * G) * SLICE + lane] = p < len ? __ldg(vals + rs + p) : T(0);
}

// y[r] = alpha * sum_p v[r,p] x[c[r,p]] + beta * y[r]; products and sums individually rounded, nonzero order
// (the CSR kernel's arithmetic for a row that lies inside one tile).
template <class T, bool HAS_REM>
__global__ void __launch_bounds__(256) spmv_sell_kernel(int nrows, int nslices, const int64_t* __restrict__ slice_off, const int* __restrict__ sinds,
                                                         const T* __restrict__ svals, const T* __restrict__ x, T alpha, T beta, const T* y_in,
                                                         T* y_out, float* out32, const T* __restrict__ rowscale, const int* __restrict__ slice_list) {
    constexpr int G = Grp<T>::G;
    using IV = typename Grp<T>::IV;
    using VV = typename Grp<T>::VV;
    pdl_trigger();
    pdl_wait();
    const int ws = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (ws >= nslices) return;
    const int s = slice_list ? __ldg(slice_list + ws) : ws;
    const int64_t off = __ldg(slice_off + s);
    const int L = (int)((__ldg(slice_off + s + 1) - off) / SLICE);
    const int ng = L / G;
    const IV* ip = reinterpret_cast<const IV*>(sinds + off) + lane;
    const VV* vp = reinterpret_cast<const VV*>(svals + off) + lane;
    T sum = T(0);
    int g = 0;
    // two groups per step: 4 streaming loads in flight, then 2 G gathers
    for (; g + 2 <= ng; g += 2) {
        const IV c0 = ldg_idx(ip + (size_t)g * SLICE), c1 = ldg_idx(ip + (size_t)(g + 1) * SLICE);
        const VV v0 = ldg_stream(vp + (size_t)g * SLICE), v1 = ldg_stream(vp + (size_t)(g + 1) * SLICE);
        const int* pc0 = reinterpret_cast<const int*>(&c0);
        const int* pc1 = reinterpret_cast<const int*>(&c1);
        const T* pv0 = reinterpret_cast<const T*>(&v0);
        const T* pv1 = reinterpret_cast<const T*>(&v1);
        T xv[2 * G];
#pragma unroll
        for (int q = 0; q < G; ++q) { xv[q] = __ldg(x + pc0[q]); xv[G + q] = __ldg(x + pc1[q]); }
#pragma unroll
        for (int q = 0; q < G; ++q) sum = add_rn(sum, mul_rn(pv0[q], xv[q]));
#pragma unroll
        for (int q = 0; q < G; ++q) sum = add_rn(sum, mul_rn(pv1[q], xv[G + q]));
    }
    if (g < ng) {
        const IV c0 = ldg_idx(ip + (size_t)g * SLICE);
        const VV v0 = ldg_stream(vp + (size_t)g * SLICE);
        const int* pc0 = reinterpret_cast<const int*>(&c0);
        const T* pv0 = reinterpret_cast<const T*>(&v0);
#pragma unroll
        for (int q = 0; q < G; ++q) sum = add_rn(sum, mul_rn(pv0[q], __ldg(x + pc0[q])));
    }
    // the L % G trailing positions, one element per lane (compiled out for plans whose slices are all whole groups)
    if (HAS_REM) {
        const int* it = sinds + off + (int64_t)ng * SLICE * G + lane;
        const T* vt = svals + off + (int64_t)ng * SLICE * G + lane;
        for (int t = 0; t < L - ng * G; ++t) sum = add_rn(sum, mul_rn(ldg_stream(vt + t * SLICE), __ldg(x + ldg_stream(it + t * SLICE))));
    }
    const int r = s * SLICE + lane;
    if (r < nrows) {
        T v = (beta == T(0)) ? alpha * sum : fma(alpha, sum, beta * y_in[r]);
        if (rowscale) {   // Jacobi: gdmv(1, d, v, 0, v), rounding sequence of the stand-alone kernel (kernels.hpp:143-145)
            const T d = __ldg(rowscale + r);
            v = add_rn(mul_rn(T(0), v), mul_rn(mul_rn(T(1), d), v));
        }
        if (y_out) y_out[r] = v;
        if (out32) out32[r] = (float)v;
    }
}

}  // namespace

struct mpg_sell_plan {
    int G = 0;
    int nslices = 0;
    int64_t total = 0;             // padded element count
    int64_t* slice_off = nullptr;  // [nslices + 1]
    int* sinds = nullptr;          // [total]
    int* slice_list = nullptr;     // partitioned matrices: slices without halo columns first
    int n_interior = 0;
    int has_rem = 0;               // some slice length is not a multiple of G
    unsigned long long uid = 0;    // distinguishes plans that happen to be allocated at the same address
};

struct mpg_packed {
    const mpg_csr* A = nullptr;
    const mpg_sell_plan* plan = nullptr;
    unsigned long long plan_uid = 0;
    int tsize = 0;
    void* svals = nullptr;
    int device = 0;
};

namespace mpg {

void sell_plan_free(mpg_sell_plan* p) {
    if (!p) return;
    cudaFree(p->slice_off); cudaFree(p->sinds); cudaFree(p->slice_list);
    delete p;
}

// build (or fetch) the packed structure for group size G; *out = nullptr if the matrix does not pack well
int sell_plan_get(mpg_ctx* ctx, const mpg_csr* A, int G, const mpg_sell_plan** out) {
    *out = nullptr;
    mpg_csr* Am = const_cast<mpg_csr*>(A);
    const int slot = (G == 4) ? 0 : 1;
    if (Am->sell_tried[slot]) { *out = Am->sell[slot]; return MPG_OK; }
    Am->sell_tried[slot] = 1;
    if (A->nrows == 0 || A->nnz == 0) return MPG_OK;
    static unsigned long long next_uid = 0;
    mpg_sell_plan* p = new mpg_sell_plan();
    p->uid = ++next_uid;
    p->G = G;
    p->nslices = (int)cdiv(A->nrows, SLICE);
    int64_t* len = nullptr;
    MPG_CUDA(ctx, pool_alloc(ctx, &len, sizeof(int64_t) * (size_t)(p->nslices + 2)));
    MPG_CUDA(ctx, pool_alloc(ctx, &p->slice_off, sizeof(int64_t) * (size_t)(p->nslices + 1)));
    MPG_CUDA(ctx, cudaMemsetAsync(len, 0, sizeof(int64_t) * (size_t)(p->nslices + 2), ctx->stream));
    const int wgrid = (int)cdiv((int64_t)p->nslices * 32, 256);
    // len[nslices] (the scan's total slot, zero) doubles as the has_rem flag until the scan has consumed it... keep it separate:
    int* rem_flag = reinterpret_cast<int*>(len + p->nslices + 1);
    sell_len_kernel<<<wgrid, 256, 0, ctx->stream>>>(A->nrows, p->nslices, A->row_map, G, len, rem_flag);
    MPG_CHECK_LAUNCH(ctx);
    sell_scan_kernel<<<1, 1024, 0, ctx->stream>>>(p->nslices, len, p->slice_off);
    MPG_CHECK_LAUNCH(ctx);
    MPG_CUDA(ctx, cudaMemcpyAsync(&p->total, p->slice_off + p->nslices, sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    MPG_CUDA(ctx, cudaMemcpyAsync(&p->has_rem, rem_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(len);
    if ((double)p->total > 1.25 * (double)A->nnz + 4096.0 || p->total >= (int64_t)1 << 40) {   // too much padding: keep CSR
        sell_plan_free(p);
        return MPG_OK;
    }
    MPG_CUDA(ctx, pool_alloc(ctx, &p->sinds, sizeof(int) * (size_t)std::max<int64_t>(p->total, 1)));
    const bool slab = A->ncols > A->nrows;
    if (slab) MPG_CUDA(ctx, pool_alloc(ctx, &p->slice_list, sizeof(int) * (size_t)p->nslices));
    if (G == 4) sell_fill_inds_kernel<4><<<wgrid, 256, 0, ctx->stream>>>(A->nrows, p->nslices, A->row_map, A->inds, p->slice_off, p->sinds, p->slice_list);
    else sell_fill_inds_kernel<2><<<wgrid, 256, 0, ctx->stream>>>(A->nrows, p->nslices, A->row_map, A->inds, p->slice_off, p->sinds, p->slice_list);
    MPG_CHECK_LAUNCH(ctx);
    if (slab) {
        // local slab of a partitioned matrix: order the slices [no halo column | some halo column]
        std::vector<int> flag((size_t)p->nslices), list((size_t)p->nslices);
        MPG_CUDA(ctx, cudaMemcpyAsync(flag.data(), p->slice_list, sizeof(int) * flag.size(), cudaMemcpyDeviceToHost, ctx->stream));
        MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        int ni = 0;
        for (int s = 0; s < p->nslices; ++s) if (!flag[(size_t)s]) list[(size_t)ni++] = s;
        p->n_interior = ni;
        for (int s = 0; s < p->nslices; ++s) if (flag[(size_t)s]) list[(size_t)ni++] = s;
        MPG_CUDA(ctx, cudaMemcpyAsync(p->slice_list, list.data(), sizeof(int) * list.size(), cudaMemcpyHostToDevice, ctx->stream));
        MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    Am->sell[slot] = p;
    *out = p;
    return MPG_OK;
}

template <class T>
int pack_update(mpg_ctx* ctx, mpg_packed* P, const T* vals) {
    const mpg_sell_plan* p = P->plan;
    const mpg_csr* A = P->A;
    ProfScope prof(ctx, MPG_PROF_ELEMENTWISE, (double)A->nnz * sizeof(T) + (double)p->total * sizeof(T));
    sell_fill_vals_kernel<T><<<(int)cdiv((int64_t)p->nslices * 32, 256), 256, 0, ctx->stream>>>(A->nrows, p->nslices, A->row_map, vals, p->slice_off,
                                                                                             static_cast<T*>(P->svals));
    MPG_CHECK_LAUNCH(ctx);
    return MPG_OK;
}
template int pack_update<float>(mpg_ctx*, mpg_packed*, const float*);
template int pack_update<double>(mpg_ctx*, mpg_packed*, const double*);

// *out = nullptr (and MPG_OK) when the structure does not pack well
template <class T>
int pack_create(mpg_ctx* ctx, const mpg_csr* A, const T* vals, mpg_packed** out) {
    *out = nullptr;
    const mpg_sell_plan* plan = nullptr;
    MPG_TRY(sell_plan_get(ctx, A, Grp<T>::G, &plan));
    if (!plan) return MPG_OK;
    mpg_packed* P = new mpg_packed();
    P->A = A; P->plan = plan; P->plan_uid = plan->uid; P->tsize = (int)sizeof(T); P->device = ctx->device;
    MPG_CUDA(ctx, pool_alloc(ctx, &P->svals, sizeof(T) * (size_t)std::max<int64_t>(plan->total, 1)));
    *out = P;
    return pack_update<T>(ctx, P, vals);
}
template int pack_create<float>(mpg_ctx*, const mpg_csr*, const float*, mpg_packed**);
template int pack_create<double>(mpg_ctx*, const mpg_csr*, const double*, mpg_packed**);

void pack_free(mpg_packed* P) {
    if (!P) return;
    cudaFree(P->svals);
    delete P;
}

bool pack_matches(const mpg_packed* P, const mpg_csr* A, int tsize) {
    const mpg_sell_plan* cur = A->sell[tsize == 4 ? 0 : 1];
    return P && P->A == A && P->tsize == tsize && cur && cur == P->plan && cur->uid == P->plan_uid;
}

template <class T>
int spmv_packed(mpg_ctx* ctx, const mpg_packed* P, T alpha, const T* x, T beta, const T* y_in, T* y_out, float* out32, const T* rowscale, int part) {
    const mpg_sell_plan* p = P->plan;
    const mpg_csr* A = P->A;
    if (part != SPMV_ALL && !p->slice_list) {
        if (part == SPMV_INTERIOR) return MPG_OK;
        part = SPMV_ALL;
    }
    const int s_first = part == SPMV_BOUNDARY ? p->n_interior : 0;
    const int s_count = part == SPMV_INTERIOR ? p->n_interior : p->nslices - s_first;
    if (s_count == 0) return MPG_OK;
    const int* list = part == SPMV_ALL ? nullptr : p->slice_list + s_first;
    // algorithmic bytes: the CSR figure of SURVEY.md §8d (the padding the packed layout reads on top is not counted)
    const double n_ = A->nrows, s_ = sizeof(T);
    const double bytes = (double)A->nnz * (s_ + 4) + 4 * (n_ + 1) + n_ * s_ + (y_out ? n_ * s_ : 0) + (beta != T(0) ? n_ * s_ : 0) + (out32 ? 4 * n_ : 0) +
                         (rowscale ? n_ * s_ : 0);
    ProfScope prof(ctx, sizeof(T) == 4 ? MPG_PROF_SPMV_F32 : MPG_PROF_SPMV_F64, bytes * ((double)s_count / p->nslices));
    const int grid = (int)cdiv((int64_t)s_count * 32, 256);
    auto kern = p->has_rem ? spmv_sell_kernel<T, true> : spmv_sell_kernel<T, false>;
    MPG_CUDA(ctx, launch_pdl(ctx, (int64_t)A->nrows, kern, grid, 256, 0, A->nrows, s_count, (const int64_t*)p->slice_off, (const int*)p->sinds, static_cast<const T*>(P->svals), x, alpha,
                             beta, y_in, y_out, out32, rowscale, list));
    MPG_CHECK_LAUNCH(ctx);
    return MPG_OK;
}
template int spmv_packed<float>(mpg_ctx*, const mpg_packed*, float, const float*, float, const float*, float*, float*, const float*, int);
template int spmv_packed<double>(mpg_ctx*, const mpg_packed*, double, const double*, double, const double*, double*, float*, const double*, int);

}  // namespace mpg

// ---- C ABI -------------------------------------------------------------------------------------------------
#define MPG_DEF_PACK(SFX, T)                                                                                                       \
    extern "C" int mpg_pack_create_##SFX(mpg_ctx* ctx, const mpg_csr* A, const T* vals, mpg_packed** out) {                       \
        MPG_REQUIRE(ctx, A && out && (vals || A->nnz == 0), "pack_create: null argument");                                        \
        return mpg::pack_create<T>(ctx, A, vals, out);                                                                            \
    }                                                                                                                              \
    extern "C" int mpg_pack_update_##SFX(mpg_ctx* ctx, mpg_packed* P, const T* vals) {                                            \
        MPG_REQUIRE(ctx, P && vals && P->tsize == (int)sizeof(T), "pack_update: null argument or wrong precision");               \
        return mpg::pack_update<T>(ctx, P, vals);                                                                                 \
    }                                                                                                                              \
    extern "C" int mpg_spmv_packed_##SFX(mpg_ctx* ctx, const mpg_packed* P, T alpha, const T* x, T beta, T* y) {                  \
        MPG_REQUIRE(ctx, P && x && y && P->tsize == (int)sizeof(T), "spmv_packed: null argument or wrong precision");             \
        return mpg::spmv_packed<T>(ctx, P, alpha, x, beta, y, y, nullptr, nullptr, mpg::SPMV_ALL);                                \
    }
MPG_DEF_PACK(f32, float)
MPG_DEF_PACK(f64, double)

extern "C" int mpg_pack_describe(const mpg_packed* P, int* group, int* nslices, int64_t* total, const int64_t** slice_off, const int** inds,
                                 const void** vals) {
    if (!P || !P->plan) return MPG_ERR_ARG;
    if (group) *group = P->plan->G;
    if (nslices) *nslices = P->plan->nslices;
    if (total) *total = P->plan->total;
    if (slice_off) *slice_off = P->plan->slice_off;
    if (inds) *inds = P->plan->sinds;
    if (vals) *vals = P->svals;
    return MPG_OK;
}

extern "C" int mpg_pack_destroy(mpg_packed* P) {
    if (!P) return MPG_OK;
    cudaSetDevice(P->device);
    mpg::pack_free(P);
    return MPG_OK;
}
