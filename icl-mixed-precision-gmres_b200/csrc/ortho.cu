// ortho.cu — tall-skinny gemv pair and the fused Arnoldi orthogonalisation (GS::add_vector).
// Reference surface: gemv kernels.hpp:118-125 (cublas?gemv, kernels_cuda.cpp:499-535); the callers are
// CGS_Kernel / MGS_Kernel / CGSR_Kernel<...,2> and GS::add_vector, Orthogonalization.hpp:51-60,76-136.
//
// The basis V is n x (k+1), column-major, column stride ldv: k+1 independent contiguous streams.  All of
// this is HBM-bound (0.5 flop/B), so the design goal is: touch every basis element once per pass, keep
// the number of passes minimal, never round-trip through the host.
//
// CGS2 as the reference formulates it is 4 gemv passes (h = V'w; w -= Vh; c = V'w; w -= Vc).  Because
// w' = w - V h is row-local, the second and third can share one read of V:
//     pass A  vpass(h_in = none):  h = V' w
//     pass B  vpass(h_in = h)   :  w' = w - V h,  c = V' w'        (V tile staged once in shared memory)
//     pass C  gemvn(+norm)      :  w'' = w' - V c, ||w''||^2
// = 3 passes with the same arithmetic per element (sequential-j fma for the updates).  The V tile
// [TR rows x k1 columns] is brought into shared memory by k1 one-dimensional TMA bulk copies
// (cp.async.bulk.shared::cluster.global + mbarrier complete_tx), multi-stage, so the loads are in flight
// while the previous tile is being consumed; each element is read from HBM once and from shared memory
// twice.  (Since round 1b: one 2-D TMA tensor load per tile issued by a dedicated producer warp.)
// Three kernels implement passes A/B, chosen by the basis width k1 only (tools/tune.py, profiles/r01b_tune_vdirect.txt):
//   vdirect  k1 <= 8          no staging, 16-byte loads straight into registers, one accumulator per column per thread
//   vrow     9 .. 32 (A) / 56 (B)   TMA-staged tiles, ROW-owner threads (no barrier inside a tile)
//   vpass    wider            TMA-staged tiles, update by row then dot products by COLUMN-owner warps
// Partial dot products stay in registers across all tiles of a CTA (warp w owns columns
// w, w+NW, ...), are written once per CTA as doubles, and the last CTA to finish sums them in a fixed
// order (deterministic, no float atomics) and runs the tiny epilogue (h += c, ...).
#include <cuda.h>

#include "common.cuh"

using namespace mpg;

namespace mpg {
int scal_devp(mpg_ctx* ctx, int64_t n, const float* a, const float* x, float* y);
int scal_devp(mpg_ctx* ctx, int64_t n, const double* a, const double* x, double* y);
int naxpy_devp(mpg_ctx* ctx, int64_t n, const float* a, const float* x, float* y);
int naxpy_devp(mpg_ctx* ctx, int64_t n, const double* a, const double* x, double* y);
int dot_dev(mpg_ctx* ctx, int64_t n, const float* x, const float* y, float* out);
int dot_dev(mpg_ctx* ctx, int64_t n, const double* x, const double* y, double* out);
int axpy_host(mpg_ctx* ctx, int64_t n, float a, const float* x, float* y);
int axpy_host(mpg_ctx* ctx, int64_t n, double a, const double* x, double* y);
template <class T> int mgs_step(mpg_ctx*, int64_t, const T*, const T*, T*, const T*, T*);
}  // namespace mpg

namespace {

// ---- mbarrier / TMA bulk helpers (sm_90+ PTX; on sm_100a these lower to SYNCS.* / UBLKCP) --------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}

enum { FIN_COEF = 0, FIN_COEF_ACCUM = 1 };

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ---------------------------------------------------------------------------------------------------------
// vpass: optional row-local update w' = w - V h_in, then partial c = V' w'   (one read of V from HBM).
//
// Warp-specialised: CT = 2*TR compute threads + one producer warp.  The producer brings a whole
// [TR rows x k1 columns] tile of V into shared memory with ONE 2-D TMA tensor load (cp.async.bulk.tensor.2d,
// box = TR x k1; rows past n are zero-filled by the hardware, so there is no tail handling) plus a 1-D tensor
// load for the w tile, signalling full[s]; it re-fills a stage as soon as the consumers release it through
// empty[s].  Loads therefore stay in flight while the consumers work on the other stage(s).
//   phase 1 (only with h_in): the two halves of the compute threads each take half of the columns for all TR
//            rows (warp-uniform split -> conflict-free shared memory reads), partial sums are combined;
//   phase 2: warp q owns columns q, q+NW, ...; lanes own rows lane+32i; accumulators live in registers
//            across all tiles of the CTA.
// The last CTA to finish sums the per-CTA partials in a fixed order and runs the epilogue
//   coef_out[j] = c_j ;  hcol[j] = c_j  (FIN_COEF)  or  hcol[j] += c_j  (FIN_COEF_ACCUM, Orthogonalization.hpp:133).
// ---------------------------------------------------------------------------------------------------------
template <class T, int TR, int MAXJ>
__global__ void __launch_bounds__(2 * TR + 32, 1)
vpass_kernel(const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapW, int64_t n, int k1, T* w, const T* h_in,
             int stages, int reverse, double* partials, int ldp, unsigned int* ticket, Epi epi, unsigned long long* dbg) {
    constexpr int CT = 2 * TR;         // compute threads
    constexpr int NW = CT / 32;        // compute warps
    constexpr int RPL = TR / 32;       // rows per lane in the dot phase
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const size_t stage_elems = (size_t)(k1 + 2) * TR;  // k1 columns of V, the w tile, the phase-1 partial
    T* tiles = reinterpret_cast<T*>(smem_raw);
    T* h_s = tiles + stage_elems * stages;
    uint64_t* full = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(h_s + ((k1 + 3) & ~3) + 4) + 7) & ~uintptr_t(7));
    uint64_t* empty = full + stages;

    const int64_t ntiles = (n + TR - 1) / TR;
    const int64_t my_count = (ntiles > blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto tile_row0 = [&](int64_t it) -> int64_t {
        int64_t tix = blockIdx.x + it * gridDim.x;
        if (reverse) tix = ntiles - 1 - tix;
        return tix * TR;
    };

    pdl_trigger_early(n);
    if (tid == CT) {   // the descriptors are kernel parameters, not data of the previous kernel: fetch them before the dependency wait
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapV) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
    }
    pdl_wait();
    if (dbg && tid == 0) dbg[blockIdx.x * 8 + 0] = globaltimer_ns();
    if (h_in) for (int j = tid; j < k1; j += CT + 32) h_s[j] = ld_fresh(h_in + j);
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        fence_mbar_init();
    }
    __syncthreads();

    T acc[MAXJ];
#pragma unroll
    for (int jj = 0; jj < MAXJ; ++jj) acc[jj] = T(0);

    if (wid == NW) {
        // ===== producer warp =====
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)((size_t)(k1 + 1) * TR * sizeof(T));
            for (int64_t it = 0; it < my_count; ++it) {
                const int s = (int)(it % stages);
                if (it >= stages) mbar_wait(empty + s, (uint32_t)(((it / stages) - 1) & 1));
                T* st = tiles + stage_elems * s;
                const int row0 = (int)tile_row0(it);
                mbar_arrive_expect_tx(full + s, bytes);
                tma_load_2d(st, &mapV, row0, 0, full + s);
                tma_load_2d(st + (size_t)k1 * TR, &mapW, row0, 0, full + s);
            }
        }
    } else {
        // ===== consumers =====
        const int half = tid / TR;          // warp-uniform: which half of the columns in phase 1
        const int row = tid - half * TR;
        const int jsplit = (k1 + 1) / 2;
        for (int64_t it = 0; it < my_count; ++it) {
            const int s = (int)(it % stages);
            const int64_t row0 = tile_row0(it);
            const int rows = (int)min((int64_t)TR, n - row0);
            T* st = tiles + stage_elems * s;
            T* ws = st + (size_t)k1 * TR;
            T* ps = ws + TR;
            mbar_wait(full + s, (uint32_t)((it / stages) & 1));
            if (dbg && tid == 0 && it == 0) dbg[blockIdx.x * 8 + 1] = globaltimer_ns();
            if (h_in) {
                // ---- phase 1: w' = w - V h ----
                const int jb = half ? jsplit : 0, je = half ? k1 : jsplit;
                T a = half ? T(0) : ws[row];
                const T* col = st + row;
#pragma unroll 8
                for (int j = jb; j < je; ++j) a = fma(-h_s[j], col[(size_t)j * TR], a);
                if (half) ps[row] = a;
                named_bar_sync(1, CT);
                if (!half) {
                    a += ps[row];
                    ws[row] = a;
                    if (row < rows) w[row0 + row] = a;
                }
                named_bar_sync(1, CT);
            }
            // ---- phase 2: c_j += sum_rows V[row, j] * w'[row] (rows past n are zero in both operands) ----
            T wl[RPL];
#pragma unroll
            for (int i = 0; i < RPL; ++i) wl[i] = ws[lane + 32 * i];
#pragma unroll
            for (int jj = 0; jj < MAXJ; ++jj) {
                const int j = wid + NW * jj;
                if (j < k1) {
                    const T* col = st + (size_t)j * TR + lane;
                    T a = T(0);
#pragma unroll
                    for (int i = 0; i < RPL; ++i) a = fma(col[32 * i], wl[i], a);
                    acc[jj] += a;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // order our shared-memory accesses before the next TMA write
            named_bar_sync(1, CT);
            if (tid == 0) mbar_arrive(empty + s);
        }
        if (dbg && tid == 0) dbg[blockIdx.x * 8 + 2] = globaltimer_ns();
        // ---- per-CTA partials (one owner warp per column) ----
#pragma unroll
        for (int jj = 0; jj < MAXJ; ++jj) {
            const int j = wid + NW * jj;
            if (j < k1) {
                const double sred = warp_sum((double)acc[jj]);
                if (lane == 0) partials[(size_t)blockIdx.x * ldp + j] = sred;
            }
        }
        if (dbg && tid == 0) dbg[blockIdx.x * 8 + 3] = globaltimer_ns();
    }
    pdl_trigger();   // streaming work of this CTA is done: the next kernel may be scheduled under the last-CTA reduction / cross-GPU combine
    const bool last = grid_last_block(ticket);
    if (dbg && tid == 0) dbg[blockIdx.x * 8 + 4] = globaltimer_ns();
    if (last) {
        // every tile of this CTA has been consumed: the staging buffers are free to serve as scratch
        last_block_finish<T>(epi, partials, ldp, gridDim.x, k1, reinterpret_cast<double*>(smem_raw));
        if (dbg && tid == 0) dbg[blockIdx.x * 8 + 5] = globaltimer_ns();
    }
}

// ---------------------------------------------------------------------------------------------------------
// gemvn: y = beta*y + sum_j (alpha*x[j]) M[:,j]   (sequential j, fma) with optional fused outputs:
//   NORM     accumulate ||y_new||^2; last CTA writes nrm = sqrt(sum) to norm_out[0] and 1/nrm to inv_out[0]
//   XUPD     mixed solution update: x64 += (double) y_new (Orthogonalization.hpp:67-73); y_new is also stored
// VEC rows per thread via 16-byte loads (VEC = 1: any alignment).
// ---------------------------------------------------------------------------------------------------------
template <class T, int VEC, bool NORM, bool XUPD>
__global__ void __launch_bounds__(256) gemvn_kernel(int64_t n, int k1, const T* __restrict__ M, int64_t ld, T alpha, const T* x, T beta,
                                                     T* y, double* x64, double* partials, unsigned int* ticket, Epi epi) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    T* ax = reinterpret_cast<T*>(smem_raw);
    pdl_trigger_early(n);
    pdl_wait();
    for (int j = threadIdx.x; j < k1; j += blockDim.x) ax[j] = alpha * ld_fresh(x + j);
    __syncthreads();
    using V4 = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
    constexpr int UN = 8;
    const int64_t nv = n / VEC;
    double nsum = 0.0;
    for (int64_t iv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; iv < nv; iv += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = iv * VEC;
        T acc[VEC];
        if (beta == T(0)) {
#pragma unroll
            for (int c = 0; c < VEC; ++c) acc[c] = T(0);
        } else if (VEC > 1) {
            const V4 t = *reinterpret_cast<const V4*>(y + i);
            const T* pt = reinterpret_cast<const T*>(&t);
#pragma unroll
            for (int c = 0; c < VEC; ++c) acc[c] = beta * pt[c];
        } else {
            acc[0] = beta * y[i];
        }
        int j = 0;
        for (; j + UN <= k1; j += UN) {
            if (VEC > 1) {
                V4 buf[UN];
#pragma unroll
                for (int u = 0; u < UN; ++u) buf[u] = ldg_stream(reinterpret_cast<const V4*>(M + (size_t)(j + u) * ld + i));
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    const T* pb = reinterpret_cast<const T*>(&buf[u]);
                    const T a = ax[j + u];
#pragma unroll
                    for (int c = 0; c < VEC; ++c) acc[c] = fma(a, pb[c], acc[c]);
                }
            } else {
                T buf[UN];
#pragma unroll
                for (int u = 0; u < UN; ++u) buf[u] = ldg_stream(M + (size_t)(j + u) * ld + i);
#pragma unroll
                for (int u = 0; u < UN; ++u) acc[0] = fma(ax[j + u], buf[u], acc[0]);
            }
        }
        for (; j < k1; ++j) {
            if (VEC > 1) {
                const V4 b = ldg_stream(reinterpret_cast<const V4*>(M + (size_t)j * ld + i));
                const T* pb = reinterpret_cast<const T*>(&b);
                const T a = ax[j];
#pragma unroll
                for (int c = 0; c < VEC; ++c) acc[c] = fma(a, pb[c], acc[c]);
            } else {
                acc[0] = fma(ax[j], ldg_stream(M + (size_t)j * ld + i), acc[0]);
            }
        }
        if (VEC > 1) {
            V4 o;
            T* po = reinterpret_cast<T*>(&o);
#pragma unroll
            for (int c = 0; c < VEC; ++c) po[c] = acc[c];
            *reinterpret_cast<V4*>(y + i) = o;
        } else {
            y[i] = acc[0];
        }
        if (XUPD) {
#pragma unroll
            for (int c = 0; c < VEC; ++c) x64[i + c] = fma(1.0, (double)acc[c], x64[i + c]);   // copy(cast) + axpy(1.0, x_temp, x)
        }
        if (NORM) {
            T q = T(0);
#pragma unroll
            for (int c = 0; c < VEC; ++c) q = fma(acc[c], acc[c], q);
            nsum += (double)q;
        }
    }
    // scalar tail rows (n % VEC)
    if (VEC > 1 && blockIdx.x == 0 && threadIdx.x < (int)(n - nv * VEC)) {
        const int64_t i = nv * VEC + threadIdx.x;
        T a = (beta == T(0)) ? T(0) : beta * y[i];
        for (int j = 0; j < k1; ++j) a = fma(ax[j], M[(size_t)j * ld + i], a);
        y[i] = a;
        if (XUPD) x64[i] = fma(1.0, (double)a, x64[i]);
        if (NORM) nsum += (double)(a * a);
    }
    pdl_trigger();
    if (NORM) {
        nsum = warp_sum(nsum);
        __shared__ double wsum[8];
        if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = nsum;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t += wsum[w];
            partials[blockIdx.x] = t;
        }
        if (grid_last_block(ticket)) last_block_finish<T>(epi, partials, 1, gridDim.x, 1);   // h(k+1,k) = nrm2(w), 1/h_final in Type  Orthogonalization.hpp:55,59
    }
}

// ---------------------------------------------------------------------------------------------------------
// generic gemv-T for the operator surface: y[j] = alpha * sum_i M[i,j] x[i] + beta*y[j], any ld / alignment.
// NC columns per sweep with register accumulators; coalesced scalar loads; x re-read per sweep (L2).
// ---------------------------------------------------------------------------------------------------------
template <class T, int NC>
__global__ void __launch_bounds__(256) gemvt_kernel(int64_t n, int ncols, const T* __restrict__ M, int64_t ld, T alpha, const T* __restrict__ x,
                                                     T beta, T* y, double* partials, int ldp, unsigned int* ticket, Epi epi) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __shared__ double red[8][NC];
    for (int j0 = 0; j0 < ncols; j0 += NC) {
        T acc[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[c] = T(0);
        const int nc = min(NC, ncols - j0);
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
            const T xi = __ldg(x + i);
            T v[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) v[c] = (c < nc) ? ldg_stream(M + (size_t)(j0 + c) * ld + i) : T(0);
#pragma unroll
            for (int c = 0; c < NC; ++c) acc[c] = fma(v[c], xi, acc[c]);
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const double s = warp_sum((double)acc[c]);
            if (lane == 0) red[wid][c] = s;
        }
        __syncthreads();
        if (threadIdx.x < nc) {
            double t = 0.0;
            for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
            partials[(size_t)blockIdx.x * ldp + j0 + threadIdx.x] = t;
        }
        __syncthreads();
    }
    if (grid_last_block(ticket)) {
        __shared__ double fin_s[256];
        last_block_finish<T>(epi, partials, ldp, gridDim.x, ncols, fin_s);
    }
}

// ---------------------------------------------------------------------------------------------------------
// gemvt_rb: h[j] = sum_i M[i,j] x[i] for tall-skinny column-major M, register accumulators, no staging.
// A CTA owns whole row blocks of RB rows; for each block it sweeps the columns NC at a time: NC independent
// 16-byte streaming loads in flight per thread, x (RB*s bytes) is re-read per sweep from L1, never from HBM.
// One partial per (row block, column) -> the last CTA adds them in block order (deterministic).
// Requires 16-byte aligned M, x and ld*sizeof(T) % 16 == 0; rows beyond n in the last vector are masked.
// ---------------------------------------------------------------------------------------------------------
template <class T, int NC>
__global__ void __launch_bounds__(256, 2) gemvt_rb_kernel(int64_t n, int ncols, const T* __restrict__ M, int64_t ld, const T* __restrict__ x,
                                                           int64_t rows_per_block, int nblocks, double* partials, int ldp, unsigned int* ticket,
                                                           Epi epi) {
    constexpr int VEC = 16 / sizeof(T);
    using V4 = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __shared__ double red[8][NC];
    for (int rb = blockIdx.x; rb < nblocks; rb += gridDim.x) {
        const int64_t r0 = (int64_t)rb * rows_per_block;
        const int64_t r1 = min(n, r0 + rows_per_block);
        for (int j0 = 0; j0 < ncols; j0 += NC) {
            const int nc = min(NC, ncols - j0);
            T acc[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) acc[c] = T(0);
            for (int64_t i = r0 + (int64_t)threadIdx.x * VEC; i < r1; i += 256 * VEC) {
                T xv[VEC];
                if (i + VEC <= n) {
                    const V4 t = *reinterpret_cast<const V4*>(x + i);
                    const T* pt = reinterpret_cast<const T*>(&t);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) xv[e] = pt[e];
                } else {
#pragma unroll
                    for (int e = 0; e < VEC; ++e) xv[e] = (i + e < n) ? x[i + e] : T(0);
                }
                V4 buf[NC];
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    if (c < nc) buf[c] = ldg_stream(reinterpret_cast<const V4*>(M + (size_t)(j0 + c) * ld + i));
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    if (c < nc) {
                        const T* pb = reinterpret_cast<const T*>(&buf[c]);
#pragma unroll
                        for (int e = 0; e < VEC; ++e) acc[c] = (i + e < n) ? fma(pb[e], xv[e], acc[c]) : acc[c];
                    }
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const double sred = warp_sum((double)acc[c]);
                if (lane == 0) red[wid][c] = sred;
            }
            __syncthreads();
            if (threadIdx.x < nc) {
                double t = 0.0;
#pragma unroll
                for (int w8 = 0; w8 < 8; ++w8) t += red[w8][threadIdx.x];
                partials[(size_t)rb * ldp + j0 + threadIdx.x] = t;
            }
            __syncthreads();
        }
    }
    if (grid_last_block(ticket)) {
        __shared__ double fin_s[256];
        last_block_finish<T>(epi, partials, ldp, nblocks, ncols, fin_s);
    }
}

// ---------------------------------------------------------------------------------------------------------
// vrow: the same operation as vpass for NARROW bases (k1 <= 32), where vpass's per-tile barriers dominate.
// Every compute thread owns RPT rows of a tile for ALL columns: it reads its V entries from shared memory once
// into registers, forms w' = w - sum_j h_j V[r,j] (sequential j), stores w', and adds V[r,j]*w'[r] into its
// k1 private accumulators - no communication between threads inside a tile, so the only synchronisation per
// tile is the full/empty mbarrier pair and the warps drift freely over the stages.  A tile is CT*RPT rows,
// brought in by the producer warp as CT*RPT/256 boxes of [256 rows x k1 columns] (+ the matching w boxes).
// After the last tile: warp xor-tree per column, cross-warp sum in warp order through shared memory, one set
// of partials per CTA, then the same last-CTA finish as vpass.
// ---------------------------------------------------------------------------------------------------------
constexpr int VROW_BOX = 256;

template <class T, int CT, int RPT, int MAXK>
__global__ void __launch_bounds__(CT + 32, 1)
vrow_kernel(const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapW, int64_t n, int k1, T* w, const T* h_in,
            int stages, int reverse, double* partials, int ldp, unsigned int* ticket, Epi epi) {
    constexpr int NW = CT / 32;
    constexpr int TRW = CT * RPT;             // rows per tile
    constexpr int NB = TRW / VROW_BOX;        // boxes per tile
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ double red[NW][MAXK];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const size_t box_elems = (size_t)(k1 + 1) * VROW_BOX;   // k1 columns of V, then the w column
    const size_t stage_elems = box_elems * NB;
    T* tiles = reinterpret_cast<T*>(smem_raw);
    T* h_s = tiles + stage_elems * stages;
    uint64_t* full = reinterpret_cast<uint64_t*>((reinterpret_cast<uintptr_t>(h_s + MAXK + 4) + 7) & ~uintptr_t(7));
    uint64_t* empty = full + stages;

    const int64_t ntiles = (n + TRW - 1) / TRW;
    const int64_t my_count = (ntiles > blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto tile_row0 = [&](int64_t it) -> int64_t {
        int64_t tix = blockIdx.x + it * gridDim.x;
        if (reverse) tix = ntiles - 1 - tix;
        return tix * TRW;
    };

    pdl_trigger_early(n);
    if (tid == CT) {   // descriptors are kernel parameters: fetch them before the dependency wait
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapV) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
    }
    pdl_wait();
    if (tid < MAXK) h_s[tid] = (h_in && tid < k1) ? ld_fresh(h_in + tid) : T(0);
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, NW); }
        fence_mbar_init();
    }
    __syncthreads();

    T acc[MAXK];
#pragma unroll
    for (int j = 0; j < MAXK; ++j) acc[j] = T(0);

    if (wid == NW) {
        // ===== producer warp =====
        if (lane == 0) {
            const uint32_t bytes = (uint32_t)(stage_elems * sizeof(T));
            for (int64_t it = 0; it < my_count; ++it) {
                const int s = (int)(it % stages);
                if (it >= stages) mbar_wait(empty + s, (uint32_t)(((it / stages) - 1) & 1));
                T* st = tiles + stage_elems * s;
                const int64_t row0 = tile_row0(it);
                mbar_arrive_expect_tx(full + s, bytes);
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    // boxes that start past n are zero-filled by the hardware like any other out-of-range element
                    const int r0 = (int)min(row0 + (int64_t)b * VROW_BOX, (int64_t)2147483000);
                    tma_load_2d(st + box_elems * b, &mapV, r0, 0, full + s);
                    tma_load_2d(st + box_elems * b + (size_t)k1 * VROW_BOX, &mapW, r0, 0, full + s);
                }
            }
        }
    } else {
        // ===== consumers: thread owns rows tid + i*CT of the tile =====
        const int b0 = tid / VROW_BOX, r = tid % VROW_BOX;
        for (int64_t it = 0; it < my_count; ++it) {
            const int s = (int)(it % stages);
            const int64_t row0 = tile_row0(it);
            const T* st = tiles + stage_elems * s;
            mbar_wait(full + s, (uint32_t)((it / stages) & 1));
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                const T* base = st + box_elems * (b0 + i * (CT / VROW_BOX)) + r;
                T v[MAXK];
#pragma unroll
                for (int j = 0; j < MAXK; ++j) v[j] = (j < k1) ? base[(size_t)j * VROW_BOX] : T(0);
                T a = base[(size_t)k1 * VROW_BOX];
                if (h_in) {
#pragma unroll
                    for (int j = 0; j < MAXK; ++j) a = fma(-h_s[j], v[j], a);      // h_s is zero past k1
                    const int64_t row = row0 + tid + (int64_t)i * CT;
                    if (row < n) w[row] = a;
                }
#pragma unroll
                for (int j = 0; j < MAXK; ++j) acc[j] = fma(v[j], a, acc[j]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
        }
#pragma unroll
        for (int j = 0; j < MAXK; ++j) {
            const double sred = warp_sum((double)acc[j]);
            if (lane == 0) red[wid][j] = sred;
        }
    }
    __syncthreads();
    if (tid < k1) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < NW; ++q) t += red[q][tid];
        partials[(size_t)blockIdx.x * ldp + tid] = t;
    }
    pdl_trigger();
    if (grid_last_block(ticket)) last_block_finish<T>(epi, partials, ldp, gridDim.x, k1, reinterpret_cast<double*>(smem_raw));
}

// ---------------------------------------------------------------------------------------------------------
// vdirect: the fused update + gemv-T with NO shared-memory staging.  A thread owns 16 bytes worth of consecutive
// rows (4 fp32 / 2 fp64) and one private accumulator per column: it loads its piece of every column straight
// from global memory into registers (coalesced 16-byte streaming loads, all k1 of them in flight together),
// forms w' = w - sum_j h_j V[r,j] (sequential j, HAS_H only), stores w', and adds V[r,j] * w'[r] into acc[j].
// Same access pattern as gemv-N, which is the fastest kernel of the path; the price is k1 accumulators (and,
// with HAS_H, the k1 loaded vectors) in registers, so it serves the widths where they fit:
//   HAS_H (pass B): k1 <= 32;   !HAS_H (pass A, h = V'w): k1 <= 104.
// After the grid-stride loop: warp xor-tree per column (fp64), cross-warp sum in warp order, per-CTA partials,
// last-CTA finish as in vpass.
// ---------------------------------------------------------------------------------------------------------
template <class T, int MAXK, bool HAS_H>
__global__ void __launch_bounds__(256) vdirect_kernel(int64_t n, int k1, const T* __restrict__ V, int64_t ldv, T* w, const T* __restrict__ h_in,
                                                      int reverse, double* partials, int ldp, unsigned int* ticket, Epi epi) {
    constexpr int VEC = 16 / sizeof(T);
    using V4 = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
    __shared__ double red[8][MAXK];
    __shared__ T h_s[MAXK];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    pdl_trigger_early(n);
    pdl_wait();
    if (HAS_H && tid < MAXK) h_s[tid] = (tid < k1) ? ld_fresh(h_in + tid) : T(0);
    __syncthreads();

    T acc[MAXK];
#pragma unroll
    for (int j = 0; j < MAXK; ++j) acc[j] = T(0);

    const int64_t nv = n / VEC;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t iv0 = (int64_t)blockIdx.x * blockDim.x + tid; iv0 < nv; iv0 += stride) {
        const int64_t i = (reverse ? nv - 1 - iv0 : iv0) * VEC;
        T a[VEC];
        {
            const V4 t = *reinterpret_cast<const V4*>(w + i);
            const T* pt = reinterpret_cast<const T*>(&t);
#pragma unroll
            for (int c = 0; c < VEC; ++c) a[c] = pt[c];
        }
        if (HAS_H) {
            V4 v[MAXK];
#pragma unroll
            for (int j = 0; j < MAXK; ++j)
                if (j < k1) v[j] = ldg_stream(reinterpret_cast<const V4*>(V + (size_t)j * ldv + i));
#pragma unroll
            for (int j = 0; j < MAXK; ++j)
                if (j < k1) {
                    const T* pv = reinterpret_cast<const T*>(&v[j]);
                    const T hj = h_s[j];
#pragma unroll
                    for (int c = 0; c < VEC; ++c) a[c] = fma(-hj, pv[c], a[c]);
                }
            V4 o;
            T* po = reinterpret_cast<T*>(&o);
#pragma unroll
            for (int c = 0; c < VEC; ++c) po[c] = a[c];
            *reinterpret_cast<V4*>(w + i) = o;
#pragma unroll
            for (int j = 0; j < MAXK; ++j)
                if (j < k1) {
                    const T* pv = reinterpret_cast<const T*>(&v[j]);
                    T q = T(0);
#pragma unroll
                    for (int c = 0; c < VEC; ++c) q = fma(pv[c], a[c], q);
                    acc[j] += q;
                }
        } else {
            constexpr int UN = 8;
#pragma unroll
            for (int j0 = 0; j0 < MAXK; j0 += UN) {
                if (j0 < k1) {
                    V4 b[UN];
#pragma unroll
                    for (int u = 0; u < UN; ++u)
                        if (j0 + u < MAXK && j0 + u < k1) b[u] = ldg_stream(reinterpret_cast<const V4*>(V + (size_t)(j0 + u) * ldv + i));
#pragma unroll
                    for (int u = 0; u < UN; ++u)
                        if (j0 + u < MAXK && j0 + u < k1) {
                            const T* pb = reinterpret_cast<const T*>(&b[u]);
                            T q = T(0);
#pragma unroll
                            for (int c = 0; c < VEC; ++c) q = fma(pb[c], a[c], q);
                            acc[j0 + u] += q;
                        }
                }
            }
        }
    }
    // rows past the last full 16-byte group: one thread each, same arithmetic
    if (blockIdx.x == 0 && tid < (int)(n - nv * VEC)) {
        const int64_t r = nv * VEC + tid;
        T a = w[r];
        if (HAS_H) {
#pragma unroll
            for (int j = 0; j < MAXK; ++j)
                if (j < k1) a = fma(-h_s[j], V[(size_t)j * ldv + r], a);
            w[r] = a;
        }
#pragma unroll
        for (int j = 0; j < MAXK; ++j)
            if (j < k1) acc[j] = fma(V[(size_t)j * ldv + r], a, acc[j]);
    }
#pragma unroll
    for (int j = 0; j < MAXK; ++j) {
        if (j < k1) {
            const double sred = warp_sum((double)acc[j]);
            if (lane == 0) red[wid][j] = sred;
        }
    }
    __syncthreads();
    if (tid < k1) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += red[q][tid];
        partials[(size_t)blockIdx.x * ldp + tid] = t;
    }
    pdl_trigger();
    if (grid_last_block(ticket)) last_block_finish<T>(epi, partials, ldp, gridDim.x, k1);
}

template <class T>
bool aligned16(const T* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

constexpr size_t kMaxDynSmem = 227 * 1024 - 4096;  // opt-in limit is 227 KB INCLUDING the kernel's static shared memory (~2.3 KB here)

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda link dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// returns 0 on success, else 1000*which + CUresult
template <class T>
int make_maps(const T* V, int64_t ldv, const T* w, int64_t n, int k1, int TR, CUtensorMap* mapV, CUtensorMap* mapW) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return -1;
    const CUtensorMapDataType dt = sizeof(T) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
    cuuint64_t gdim[2] = {(cuuint64_t)n, (cuuint64_t)k1};
    cuuint64_t gstr[1] = {(cuuint64_t)ldv * sizeof(T)};
    cuuint32_t box[2] = {(cuuint32_t)TR, (cuuint32_t)k1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(mapV, dt, 2, const_cast<T*>(V), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return 1000 + (int)r;
    // w as an n x 1 matrix (same 2-D load path; rows past n are zero-filled)
    cuuint64_t gdim1[2] = {(cuuint64_t)n, 1};
    cuuint64_t gstr1[1] = {(((cuuint64_t)n * sizeof(T) + 15) / 16) * 16};
    cuuint32_t box1[2] = {(cuuint32_t)TR, 1};
    r = enc(mapW, dt, 2, const_cast<T*>(w), gdim1, gstr1, box1, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return 2000 + (int)r;
    return 0;
}

template <class T>
struct VpassCfg;
template <> struct VpassCfg<float> { static constexpr int TR = 256; };
template <> struct VpassCfg<double> { static constexpr int TR = 128; };

template <class T>
int cached_maps(mpg_ctx* ctx, const T* V, int64_t ldv, const T* w, int64_t n, int k1, int box_rows, const CUtensorMap** mapV, const CUtensorMap** mapW);

template <class T, int TR, int MAXJ>
int launch_vpass_inst(mpg_ctx* ctx, int64_t n, int k1, const T* V, int64_t ldv, T* w, const T* h_in, int fin, T* coef_out, T* hcol) {
    const CUtensorMap *pmapV, *pmapW;
    MPG_TRY(cached_maps<T>(ctx, V, ldv, w, n, k1, TR, &pmapV, &pmapW));
    const CUtensorMap& mapV = *pmapV;
    const CUtensorMap& mapW = *pmapW;
    const size_t stage_bytes = (size_t)(k1 + 2) * TR * sizeof(T);
    const size_t extra = sizeof(T) * (size_t)(((k1 + 3) & ~3) + 4) + 8 * 2 * 8 + 32;
    int stages = (int)std::min<size_t>(8, (kMaxDynSmem - extra) / stage_bytes);
    if (ctx->tune.vpass_stages > 0) stages = std::min(stages, ctx->tune.vpass_stages);
    if (stages < 1) return fail(ctx, MPG_ERR_ARG, "vpass: tile does not fit in shared memory");
    // 544 threads x ~75-100 registers: one CTA per SM; the producer warp keeps up to `stages` tiles in flight
    const size_t smem = stage_bytes * stages + extra;
    const int ctas_per_sm = 1;
    const int64_t ntiles = cdiv(n, TR);
    const int grid = (int)std::min<int64_t>(std::min<int64_t>(ntiles, (int64_t)ctx->num_sms * ctas_per_sm), kMaxPartBlocks);
    auto kern = vpass_kernel<T, TR, MAXJ>;
    static bool attr_set[64] = {};   // per instantiation and device (the attribute is per device)
    if (!attr_set[ctx->device & 63]) {
        MPG_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxDynSmem));
        attr_set[ctx->device & 63] = true;
    }
    // algorithmic bytes (fused CGS2 = 3 k1 n s + 4 n s): pass A reads V only (w was just written by the
    // SpMV and is counted there), pass B reads V, reads w, writes w'
    ProfScope prof(ctx, MPG_PROF_VPASS, (double)k1 * n * sizeof(T) + (h_in ? 2.0 * n * sizeof(T) : 0.0));
    // traversal direction is a function of the pass, never of call history (bit-reproducible results): the plain
    // gemv-T pass runs forward, the fused update+gemv-T pass runs backward so it starts on the part of V the previous
    // pass left in L2, and the following gemv-N pass (forward) starts on the part this one leaves there
    const int reverse = (ctx->tune.vpass_serpentine && h_in) ? 1 : 0;
    const Epi epi = make_epi(ctx, fin == FIN_COEF_ACCUM ? EPI_COEF_ACCUM : EPI_COEF, coef_out, hcol, 0.0, 0.0);
    MPG_CUDA(ctx, launch_pdl(ctx, n, kern, grid, 2 * TR + 32, smem, mapV, mapW, n, k1, w, h_in, stages, reverse, ctx->partials, (int)(kMaxCols + 8), ctx->ticket, epi, ctx->dbg));
    MPG_CHECK_LAUNCH(ctx);
    return dist_finish_reduction(ctx, epi, k1, (int)sizeof(T));
}

// tensor maps are pure functions of (pointer, n, ldv, k1, box rows): keep the last few (one Arnoldi cycle re-uses one per width).
// The cache belongs to the CONTEXT (contexts on different threads / devices never share mutable state).
struct MapKey { const void* V; const void* w; int64_t n, ldv; int k1, tr, ts; };
struct MapSlot { MapKey key; CUtensorMap mapV, mapW; bool valid; };
struct MapCache { MapSlot slot[2048]; };
void map_cache_free(void* p) { delete static_cast<MapCache*>(p); }

template <class T>
int cached_maps(mpg_ctx* ctx, const T* V, int64_t ldv, const T* w, int64_t n, int k1, int box_rows, const CUtensorMap** mapV, const CUtensorMap** mapW) {
    if (!ctx->ortho_cache) {
        MapCache* c = new MapCache();
        memset(c, 0, sizeof(MapCache));
        ctx->ortho_cache = c;
        ctx->ortho_cache_free = map_cache_free;
    }
    MapCache* cache = static_cast<MapCache*>(ctx->ortho_cache);
    MapSlot& slot = cache->slot[(((size_t)k1 * 2 + (sizeof(T) == 8)) * 4 + (box_rows == VROW_BOX ? 0 : (box_rows == 256 ? 1 : 2))) & 2047];
    MapKey key;
    memset(&key, 0, sizeof(MapKey));
    key.V = V; key.w = w; key.n = n; key.ldv = ldv; key.k1 = k1; key.tr = box_rows; key.ts = (int)sizeof(T);
    if (!slot.valid || memcmp(&slot.key, &key, sizeof(MapKey)) != 0) {
        const int mrc = make_maps<T>(V, ldv, w, n, k1, box_rows, &slot.mapV, &slot.mapW);
        memcpy(&slot.key, &key, sizeof(MapKey));
        slot.valid = (mrc == 0);
        if (mrc != 0)
            return fail(ctx, MPG_ERR_CUDA, "vpass: cuTensorMapEncodeTiled failed, code " + std::to_string(mrc) + " (n=" + std::to_string(n) + " k1=" +
                                               std::to_string(k1) + " ldv=" + std::to_string(ldv) + ")");
    }
    *mapV = &slot.mapV;
    *mapW = &slot.mapW;
    return MPG_OK;
}

template <class T, int CT, int RPT, int MAXK>
int launch_vrow_inst(mpg_ctx* ctx, int64_t n, int k1, const T* V, int64_t ldv, T* w, const T* h_in, int fin, T* coef_out, T* hcol) {
    const CUtensorMap *mapV, *mapW;
    MPG_TRY(cached_maps<T>(ctx, V, ldv, w, n, k1, VROW_BOX, &mapV, &mapW));
    constexpr int TRW = CT * RPT;
    const size_t stage_bytes = (size_t)(k1 + 1) * TRW * sizeof(T);
    const size_t extra = sizeof(T) * (size_t)(MAXK + 4) + 8 * 2 * 8 + 32 + 128;
    const size_t budget = kMaxDynSmem - sizeof(double) * (CT / 32) * MAXK;   // the cross-warp buffer is static shared memory
    int stages = (int)std::min<size_t>(8, (budget - extra) / stage_bytes);
    if (ctx->tune.vpass_stages > 0) stages = std::min(stages, ctx->tune.vpass_stages);
    if (stages < 1) return fail(ctx, MPG_ERR_ARG, "vrow: tile does not fit in shared memory");
    const size_t smem = stage_bytes * stages + extra;
    const int64_t ntiles = cdiv(n, TRW);
    const int grid = (int)std::min<int64_t>(std::min<int64_t>(ntiles, (int64_t)ctx->num_sms), kMaxPartBlocks);
    auto kern = vrow_kernel<T, CT, RPT, MAXK>;
    static bool attr_set[64] = {};   // per instantiation and device (the attribute is per device)
    if (!attr_set[ctx->device & 63]) {
        MPG_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
        attr_set[ctx->device & 63] = true;
    }
    ProfScope prof(ctx, MPG_PROF_VPASS, (double)k1 * n * sizeof(T) + (h_in ? 2.0 * n * sizeof(T) : 0.0));
    const int reverse = (ctx->tune.vpass_serpentine && h_in) ? 1 : 0;   // same rule as vpass: a function of the pass only
    const Epi epi = make_epi(ctx, fin == FIN_COEF_ACCUM ? EPI_COEF_ACCUM : EPI_COEF, coef_out, hcol, 0.0, 0.0);
    MPG_CUDA(ctx, launch_pdl(ctx, n, kern, grid, CT + 32, smem, *mapV, *mapW, n, k1, w, h_in, stages, reverse, ctx->partials, (int)(kMaxCols + 8), ctx->ticket, epi));
    MPG_CHECK_LAUNCH(ctx);
    return dist_finish_reduction(ctx, epi, k1, (int)sizeof(T));
}

template <class T>
int launch_vrow(mpg_ctx* ctx, int64_t n, int k1, const T* V, int64_t ldv, T* w, const T* h_in, int fin, T* coef_out, T* hcol) {
    constexpr int CT = sizeof(T) == 4 ? 512 : 256;
    if (k1 <= 8) return launch_vrow_inst<T, CT, 4, 8>(ctx, n, k1, V, ldv, w, h_in, fin, coef_out, hcol);
    if (k1 <= 16) return launch_vrow_inst<T, CT, 2, 16>(ctx, n, k1, V, ldv, w, h_in, fin, coef_out, hcol);
    if (k1 <= 32) return launch_vrow_inst<T, CT, 1, 32>(ctx, n, k1, V, ldv, w, h_in, fin, coef_out, hcol);
    if constexpr (sizeof(T) == 4) {
        if (k1 <= 48) return launch_vrow_inst<T, 256, 1, 48>(ctx, n, k1, V, ldv, w, h_in, fin, coef_out, hcol);
        if (k1 <= 64) return launch_vrow_inst<T, 256, 1, 64>(ctx, n, k1, V, ldv, w, h_in, fin, coef_out, hcol);
    }
    return fail(ctx, MPG_ERR_ARG, "vrow: too many columns");
}

template <class T, int MAXK, bool HAS_H>
int launch_vdirect_inst(mpg_ctx* ctx, int64_t n, int k1, const T* V, int64_t ldv, T* w, const T* h_in, int fin, T* coef_out, T* hcol) {
    constexpr int VEC = 16 / sizeof(T);
    auto kern = vdirect_kernel<T, MAXK, HAS_H>;
    static int occ[64] = {};   // per instantiation and device (an int store is atomic; two contexts computing it write the same value)
    int ctas_per_sm = occ[ctx->device & 63];
    if (ctas_per_sm == 0) {
        MPG_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, 256, 0));
        ctas_per_sm = std::max(1, std::min(ctas_per_sm, 4));
        occ[ctx->device & 63] = ctas_per_sm;
    }
    const int64_t groups = std::max<int64_t>(1, n / VEC);
    int grid = (int)std::min<int64_t>(cdiv(groups, 256), (int64_t)ctx->num_sms * ctas_per_sm);
    grid = std::max(1, std::min(grid, kMaxPartBlocks));
    ProfScope prof(ctx, MPG_PROF_VPASS, (double)k1 * n * sizeof(T) + (h_in ? 2.0 * n * sizeof(T) : 0.0));
    const int reverse = (ctx->tune.vpass_serpentine && h_in) ? 1 : 0;   // same rule as vpass: a function of the pass only
    const Epi epi = make_epi(ctx, fin == FIN_COEF_ACCUM ? EPI_COEF_ACCUM : EPI_COEF, coef_out, hcol, 0.0, 0.0);
    MPG_CUDA(ctx, launch_pdl(ctx, n, kern, grid, 256, 0, n, k1, V, ldv, w, h_in, reverse, ctx->partials, (int)(kMaxCols + 8), ctx->ticket, epi));
    MPG_CHECK_LAUNCH(ctx);
    return dist_finish_reduction(ctx, epi, k1, (int)sizeof(T));
}

// widest basis the register-resident kernel is built for.  Measured (profiles/r01b_tune_vdirect.txt): it wins up to 8
// columns; from 9 on the k1 accumulators (+ k1 live 16-byte vectors in pass B) cost too much occupancy and the
// shared-memory staged kernels are faster.
constexpr int kVdirectCap = 16;

template <class T>
int launch_vdirect(mpg_ctx* ctx, int64_t n, int k1, const T* V, int64_t ldv, T* w, const T* h_in, int fin, T* coef_out, T* hcol) {
#define MPG_VD(K, H) return launch_vdirect_inst<T, K, H>(ctx, n, k1, V, ldv, w, h_in, fin, coef_out, hcol)
    if (h_in) {
        if (k1 <= 8) MPG_VD(8, true);
        if (k1 <= 16) MPG_VD(16, true);
    } else {
        if (k1 <= 8) MPG_VD(8, false);
        if (k1 <= 16) MPG_VD(16, false);
    }
#undef MPG_VD
    return fail(ctx, MPG_ERR_ARG, "vdirect: too many columns");
}

template <class T>
int launch_vpass(mpg_ctx* ctx, int64_t n, int k1, const T* V, int64_t ldv, T* w, const T* h_in, int fin, T* coef_out, T* hcol) {
    {
        const int cap = std::min(h_in ? ctx->tune.vdirect_max_cols_b : ctx->tune.vdirect_max_cols_a, kVdirectCap);
        if (k1 <= cap && (ldv * sizeof(T)) % 16 == 0) return launch_vdirect<T>(ctx, n, k1, V, ldv, w, h_in, fin, coef_out, hcol);
    }
    if (k1 <= (h_in ? ctx->tune.vrow_max_cols : std::min(ctx->tune.vrow_max_cols, ctx->tune.vrow_max_cols_a)) && k1 <= (sizeof(T) == 4 ? 64 : 32))
        return launch_vrow<T>(ctx, n, k1, V, ldv, w, h_in, fin, coef_out, hcol);
    constexpr int TR = VpassCfg<T>::TR;
    constexpr int NW = 2 * TR / 32;
    const int need = (k1 + NW - 1) / NW;
    if (need <= 2) return launch_vpass_inst<T, TR, 2>(ctx, n, k1, V, ldv, w, h_in, fin, coef_out, hcol);
    if (need <= 4) return launch_vpass_inst<T, TR, 4>(ctx, n, k1, V, ldv, w, h_in, fin, coef_out, hcol);
    if (need <= 7) return launch_vpass_inst<T, TR, 7>(ctx, n, k1, V, ldv, w, h_in, fin, coef_out, hcol);
    if (need <= 13) return launch_vpass_inst<T, TR, 13>(ctx, n, k1, V, ldv, w, h_in, fin, coef_out, hcol);
    return fail(ctx, MPG_ERR_ARG, "vpass: too many columns");
}

}  // namespace

namespace mpg {

// widest basis block the fused vpass accepts for T: two stages of [TR x (k1+2)] must fit in shared memory and each
// compute warp owns at most 13 columns
template <class T>
int vpass_max_cols() {
    constexpr int TR = VpassCfg<T>::TR;
    const int by_regs = 13 * (2 * TR / 32);
    const int by_smem = (int)((kMaxDynSmem - 4096) / 2 / (TR * sizeof(T))) - 2;
    return std::min(std::min(by_regs, by_smem), 256);
}
template int vpass_max_cols<float>();
template int vpass_max_cols<double>();

// TMA needs 16-byte aligned base addresses and a 16-byte multiple column stride; callers fall back to the
// unfused gemv-T / gemv-N kernels otherwise
template <class T>
bool vpass_ok(const T* V, int64_t ldv, const T* w, int64_t n, int k1) {
    return aligned16(V) && aligned16(w) && ((ldv * sizeof(T)) % 16 == 0) && k1 >= 1 && k1 <= vpass_max_cols<T>() && n < (int64_t)2147483647 &&
           encode_tiled() != nullptr;
}
template bool vpass_ok<float>(const float*, int64_t, const float*, int64_t, int);
template bool vpass_ok<double>(const double*, int64_t, const double*, int64_t, int);

template <class T>
int vpass(mpg_ctx* ctx, int64_t n, int k1, const T* V, int64_t ldv, T* w, const T* h_in, int accumulate, T* coef_out, T* hcol) {
    return launch_vpass<T>(ctx, n, k1, V, ldv, w, h_in, accumulate ? FIN_COEF_ACCUM : FIN_COEF, coef_out, hcol);
}
template int vpass<float>(mpg_ctx*, int64_t, int, const float*, int64_t, float*, const float*, int, float*, float*);
template int vpass<double>(mpg_ctx*, int64_t, int, const double*, int64_t, double*, const double*, int, double*, double*);

template <class T>
int gemvn(mpg_ctx* ctx, int64_t n, int k1, const T* M, int64_t ld, T alpha, const T* x, T beta, T* y, bool want_norm, T* norm_out, T* inv_out,
          double* x64) {
    if (n <= 0) return MPG_OK;
    constexpr int VEC = 16 / sizeof(T);
    const bool vec_ok = aligned16(M) && aligned16(y) && ((ld * sizeof(T)) % 16 == 0) && (!x64 || aligned16(x64));
    const size_t smem = sizeof(T) * (size_t)std::max(k1, 1);
    const int64_t work = vec_ok ? cdiv(n, VEC) : n;
    int grid = (int)std::min<int64_t>(cdiv(work, 256), (int64_t)ctx->num_sms * ctx->tune.gemvn_ctas_per_sm);
    grid = std::max(1, std::min(grid, kMaxPartBlocks));
    // gemv-N: k1 n s + 2 n s (read y, write y) [+ 16 n for the fp64 x read-modify-write of the mixed update]
    ProfScope prof(ctx, MPG_PROF_GEMVN, (double)k1 * n * sizeof(T) + (beta != T(0) ? 2.0 : 1.0) * n * sizeof(T) + (x64 ? 16.0 * n : 0.0));
    const Epi epi = want_norm ? make_epi(ctx, EPI_NORM_INV, norm_out, inv_out, 0.0, 0.0) : Epi{EPI_NORM_INV, norm_out, inv_out, 0.0, 0.0, nullptr, PeerComm()};
#define MPG_GEMVN(VV, NN, XX)                                                                                                   \
    MPG_CUDA(ctx, launch_pdl(ctx, n, gemvn_kernel<T, VV, NN, XX>, grid, 256, smem, n, k1, M, ld, alpha, x, beta, y, x64, ctx->partials, ctx->ticket, epi))
    if (vec_ok) {
        if (want_norm) MPG_GEMVN(VEC, true, false);
        else if (x64) MPG_GEMVN(VEC, false, true);
        else MPG_GEMVN(VEC, false, false);
    } else {
        if (want_norm) MPG_GEMVN(1, true, false);
        else if (x64) MPG_GEMVN(1, false, true);
        else MPG_GEMVN(1, false, false);
    }
#undef MPG_GEMVN
    MPG_CHECK_LAUNCH(ctx);
    if (want_norm) return dist_finish_reduction(ctx, epi, 1, (int)sizeof(T));
    return MPG_OK;
}
template int gemvn<float>(mpg_ctx*, int64_t, int, const float*, int64_t, float, const float*, float, float*, bool, float*, float*, double*);
template int gemvn<double>(mpg_ctx*, int64_t, int, const double*, int64_t, double, const double*, double, double*, bool, double*, double*, double*);

template <class T>
int gemvt(mpg_ctx* ctx, int64_t n, int ncols, const T* M, int64_t ld, T alpha, const T* x, T beta, T* y) {
    if (ncols <= 0) return MPG_OK;
    if (ncols > kMaxCols) return fail(ctx, MPG_ERR_ARG, "gemv-T: more than 256 columns");
    constexpr int VEC = 16 / sizeof(T);
    // fast path: 16-byte aligned columns whose vector loads cannot run past the allocation (ld >= n rounded up)
    if (ctx->tune.gemvt_rb && n >= 4096 && aligned16(M) && aligned16(x) && ((ld * sizeof(T)) % 16 == 0) && ld >= ((n + VEC - 1) / VEC) * VEC) {
        int64_t rpb = std::max<int64_t>(ctx->tune.gemvt_rows_per_block, cdiv(n, kMaxPartBlocks));
        rpb = ((rpb + 256 * VEC - 1) / (256 * VEC)) * (256 * VEC);
        const int nblocks = (int)cdiv(n, rpb);
        const int grid = std::min(nblocks, ctx->num_sms * 2);
        ProfScope prof(ctx, MPG_PROF_GEMVT, (double)ncols * n * sizeof(T) + (double)n * sizeof(T));
        constexpr int NC = sizeof(T) == 4 ? 16 : 8;
        const Epi epi = make_epi(ctx, EPI_GEMVT, y, nullptr, (double)alpha, (double)beta);
        gemvt_rb_kernel<T, NC><<<grid, 256, 0, ctx->stream>>>(n, ncols, M, ld, x, rpb, nblocks, ctx->partials, kMaxCols + 8, ctx->ticket, epi);
        MPG_CHECK_LAUNCH(ctx);
        return dist_finish_reduction(ctx, epi, ncols, (int)sizeof(T));
    }
    int grid = (int)std::min<int64_t>(std::max<int64_t>(1, cdiv(n, 256 * 4)), (int64_t)ctx->num_sms * 4);
    grid = std::min(grid, kMaxPartBlocks);
    ProfScope prof(ctx, MPG_PROF_GEMVT, (double)ncols * n * sizeof(T) + (double)n * sizeof(T));
    const Epi epi = make_epi(ctx, EPI_GEMVT, y, nullptr, (double)alpha, (double)beta);
    gemvt_kernel<T, 8><<<grid, 256, 0, ctx->stream>>>(n, ncols, M, ld, alpha, x, beta, y, ctx->partials, kMaxCols + 8, ctx->ticket, epi);
    MPG_CHECK_LAUNCH(ctx);
    return dist_finish_reduction(ctx, epi, ncols, (int)sizeof(T));
}
template int gemvt<float>(mpg_ctx*, int64_t, int, const float*, int64_t, float, const float*, float, float*);
template int gemvt<double>(mpg_ctx*, int64_t, int, const double*, int64_t, double, const double*, double, double*);

// GS::add_vector (Orthogonalization.hpp:51-60).  scratch: >= k1 + 2 elements of T in device memory
// (CGSR's `weights`, Orthogonalization.hpp:113, plus the 1/norm scalar).
template <class T>
int add_vector(mpg_ctx* ctx, int orth, int64_t n, int64_t k, T* V, int64_t ldv, T* w, T* hcol, T* scratch, bool skip_normalize) {
    const int k1 = (int)k + 1;
    T* weights = scratch;
    T* inv = scratch + k1;
    T* vnext = V + (size_t)(k + 1) * ldv;
    if (orth == MPG_ORTH_MGS) {
        // Orthogonalization.hpp:98-106: k+1 x { dot -> h(j,k) on device ; w -= h(j,k) v_j }
        if (ctx->tune.mgs_fused && aligned16(V) && aligned16(w) && (ldv * sizeof(T)) % 16 == 0) {
            // pairwise fused: the naxpy of column j and the dot of column j + 1 share one pass over w (k + 2 launches instead of
            // 2 k + 3, 4 n s instead of 5 n s bytes per column); the last naxpy carries the norm.  Bit-identical to the loop below.
            MPG_TRY(dot_dev(ctx, n, w, V, hcol));
            for (int j = 0; j + 1 < k1; ++j) MPG_TRY(mgs_step<T>(ctx, n, V + (size_t)j * ldv, V + (size_t)(j + 1) * ldv, w, hcol + j, hcol + j + 1));
            MPG_TRY(gemvn<T>(ctx, n, 1, V + (size_t)(k1 - 1) * ldv, ldv, T(-1), hcol + (k1 - 1), T(1), w, true, hcol + k1, inv, nullptr));
        } else {
            for (int j = 0; j < k1; ++j) {
                MPG_TRY(dot_dev(ctx, n, w, V + (size_t)j * ldv, hcol + j));
                MPG_TRY(naxpy_devp(ctx, n, hcol + j, V + (size_t)j * ldv, w));
            }
            // nrm2(w, h(k+1,k)) as a gemv-N with zero columns would be wasteful: reuse the NORM epilogue with k1 = 0
            MPG_TRY(gemvn<T>(ctx, n, 0, V, ldv, T(0), hcol, T(1), w, true, hcol + k1, inv, nullptr));
        }
    } else {
        // narrow bases (k1 < fuse_min_cols) are dominated by the staged kernel's per-tile cost: the register kernels
        // (gemv-T row-block sweep + gemv-N) are faster there even though they read V four times instead of three
        const bool fused = ctx->tune.cgs2_fused && k1 >= ctx->tune.fuse_min_cols && vpass_ok<T>(V, ldv, w, n, k1);
        if (fused) {
            if (ctx->tune.passA_rb) MPG_TRY(gemvt<T>(ctx, n, k1, V, ldv, T(1), w, T(0), hcol));
            else MPG_TRY(vpass<T>(ctx, n, k1, V, ldv, w, nullptr, 0, hcol, hcol));                          // h = V'w          :82,:126
            if (orth == MPG_ORTH_CGSR) {
                MPG_TRY(vpass<T>(ctx, n, k1, V, ldv, w, hcol, 1, weights, hcol));                           // w-=Vh; c=V'w; h+=c :127-133
                MPG_TRY(gemvn<T>(ctx, n, k1, V, ldv, T(-1), weights, T(1), w, true, hcol + k1, inv, nullptr));  // w-=Vc; ||w||   :131,:55
            } else {
                MPG_TRY(gemvn<T>(ctx, n, k1, V, ldv, T(-1), hcol, T(1), w, true, hcol + k1, inv, nullptr));     // w-=Vh; ||w||   :87,:55
            }
        } else {
            MPG_TRY(gemvt<T>(ctx, n, k1, V, ldv, T(1), w, T(0), hcol));
            if (orth == MPG_ORTH_CGSR) {
                MPG_TRY(gemvn<T>(ctx, n, k1, V, ldv, T(-1), hcol, T(1), w, false, nullptr, nullptr, nullptr));
                MPG_TRY(gemvt<T>(ctx, n, k1, V, ldv, T(1), w, T(0), weights));
                MPG_TRY(gemvn<T>(ctx, n, k1, V, ldv, T(-1), weights, T(1), w, true, hcol + k1, inv, nullptr));
                MPG_TRY(axpy_host(ctx, k1, T(1), weights, hcol));                                               // h += c  :133
            } else {
                MPG_TRY(gemvn<T>(ctx, n, k1, V, ldv, T(-1), hcol, T(1), w, true, hcol + k1, inv, nullptr));
            }
        }
    }
    // V[:,k+1] = w * (1/h(k+1,k))   Orthogonalization.hpp:58-59 (one pass instead of copy + scal); the solver folds this
    // pass into its arnoldi_tail launch (1/h(k+1,k) is left at scratch[k+1])
    if (skip_normalize) return MPG_OK;
    return scal_devp(ctx, n, inv, w, vnext);
}
template int add_vector<float>(mpg_ctx*, int, int64_t, int64_t, float*, int64_t, float*, float*, float*, bool);
template int add_vector<double>(mpg_ctx*, int, int64_t, int64_t, double*, int64_t, double*, double*, double*, bool);

}  // namespace mpg

// ---- C ABI -------------------------------------------------------------------------------------------------
#define MPG_DEF_GEMV(SFX, T)                                                                                                      \
    extern "C" int mpg_gemv_##SFX(mpg_ctx* ctx, int trans, int64_t nr, int64_t nc, T alpha, const T* M, int64_t ld, const T* x,   \
                                  T beta, T* y) {                                                                                 \
        MPG_REQUIRE(ctx, nr >= 0 && nc >= 0 && ld >= nr, "gemv: bad dims");                                                       \
        MPG_REQUIRE(ctx, nc <= kMaxCols, "gemv: at most 256 columns (restart length + 1)");                                       \
        if (trans) return mpg::gemvt<T>(ctx, nr, (int)nc, M, ld, alpha, x, beta, y);                                              \
        return mpg::gemvn<T>(ctx, nr, (int)nc, M, ld, alpha, x, beta, y, false, nullptr, nullptr, nullptr);                       \
    }                                                                                                                             \
    extern "C" int mpg_add_vector_##SFX(mpg_ctx* ctx, int orth, int64_t n, int64_t k, T* V, int64_t ldv, T* w, T* hcol) {         \
        MPG_REQUIRE(ctx, n >= 0 && k >= 0 && k + 2 <= kMaxCols && ldv >= n, "add_vector: bad dims");                              \
        MPG_REQUIRE(ctx, orth >= 0 && orth <= 2, "add_vector: bad orth");                                                         \
        T* scratch = reinterpret_cast<T*>(ctx->dscal + 64); /* k1 + 2 <= 264 elements */                                          \
        return mpg::add_vector<T>(ctx, orth, n, k, V, ldv, w, hcol, scratch, false);                                                     \
    }
MPG_DEF_GEMV(f32, float)
MPG_DEF_GEMV(f64, double)
