// blas1.cu — context, device memory, BLAS-1, casts, Givens / least-squares kernels.
// Reference surface: kernels.hpp:11-114,128-151; reference CUDA backend: kernels_cuda.cpp:111-494,538-572.
#include <mutex>
#include <random>
#include <unordered_map>

#include "common.cuh"

using namespace mpg;

namespace mpg {
// ---- block cache on top of the stream-ordered pool (see common.cuh) --------------------------------------------------------
namespace blockcache {
struct CachedBlock { void* p; size_t bytes; };
struct LiveBlock { size_t bytes; int dev; cudaStream_t stream; unsigned epoch; };
struct BlockCache {
    std::mutex mu;
    std::unordered_map<void*, LiveBlock> live;                   // handed out: ptr -> (bytes, device, stream it was allocated on, stream epoch)
    std::vector<CachedBlock> parked[64];                         // per device
    size_t parked_bytes[64] = {0};
    unsigned epoch = 0;                                          // bumped by mpg_ctx_set_stream / mpg_ctx_use_own_stream
};
BlockCache& block_cache() { static BlockCache c; return c; }
constexpr size_t kCacheMinBytes = (size_t)1 << 20;
constexpr size_t kCacheMaxParked = (size_t)24 << 30;             // per device
}  // namespace blockcache
using namespace blockcache;

cudaError_t pool_alloc(mpg_ctx* ctx, void** p, size_t bytes) {
    BlockCache& c = block_cache();
    const int dev = ctx->device & 63;
    if (bytes >= kCacheMinBytes) {
        std::lock_guard<std::mutex> lock(c.mu);
        auto& v = c.parked[dev];
        int best = -1;
        for (int i = 0; i < (int)v.size(); ++i)
            if (v[(size_t)i].bytes >= bytes && v[(size_t)i].bytes <= bytes + bytes / 4 + kCacheMinBytes && (best < 0 || v[(size_t)i].bytes < v[(size_t)best].bytes)) best = i;
        if (best >= 0) {
            *p = v[(size_t)best].p;
            c.live[*p] = {v[(size_t)best].bytes, dev, ctx->stream, c.epoch};
            c.parked_bytes[dev] -= v[(size_t)best].bytes;
            v.erase(v.begin() + best);
            return cudaSuccess;
        }
    }
    const cudaError_t e = cudaMallocAsync(p, bytes, ctx->stream);
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> lock(c.mu);
        c.live[*p] = {bytes, dev, ctx->stream, c.epoch};
    }
    return e;
}
// Every library allocation is only ever touched by work on its context's stream, so "nobody is using the block any more" is a
// synchronisation of THAT stream - not of the device: a device-wide wait would also wait for the copy stream of the overlapped host
// path (hostpath.cu), i.e. for gigabytes still crossing PCIe, in the middle of a solve.  If the context's stream was switched since
// the block was handed out (stream epoch), the conservative device-wide wait is kept.
void pool_free(void* p) {
    if (!p) return;
    BlockCache& c = block_cache();
    LiveBlock b{0, 0, nullptr, 0};
    bool same_epoch = false;
    {
        std::lock_guard<std::mutex> lock(c.mu);
        auto it = c.live.find(p);
        if (it != c.live.end()) { b = it->second; c.live.erase(it); same_epoch = (b.epoch == c.epoch); }
    }
    if (b.bytes == 0) { cudaFree(p); return; }                    // not ours (plain cudaMalloc)
    if (same_epoch) cudaStreamSynchronize(b.stream); else cudaDeviceSynchronize();
    if (b.bytes < kCacheMinBytes) { cudaFreeAsync(p, same_epoch ? b.stream : (cudaStream_t)0); return; }
    std::lock_guard<std::mutex> lock(c.mu);
    if (c.parked_bytes[b.dev] + b.bytes > kCacheMaxParked || c.parked[b.dev].size() >= 256) { cudaFreeAsync(p, same_epoch ? b.stream : (cudaStream_t)0); return; }
    c.parked[b.dev].push_back({p, b.bytes});
    c.parked_bytes[b.dev] += b.bytes;
}
void pool_stream_changed() {
    BlockCache& c = block_cache();
    std::lock_guard<std::mutex> lock(c.mu);
    ++c.epoch;
}
void pool_trim(int device) {
    BlockCache& c = block_cache();
    std::vector<CachedBlock> drop;
    {
        std::lock_guard<std::mutex> lock(c.mu);
        drop.swap(c.parked[device & 63]);
        c.parked_bytes[device & 63] = 0;
    }
    for (auto& b : drop) cudaFree(b.p);
}

// device-side waits on a peer GPU that timed out leave a code in the context's error word (common.cuh wait_flag)
int check_dev_err(mpg_ctx* ctx) {
    const unsigned int e = ctx->dev_err ? *reinterpret_cast<volatile unsigned int*>(ctx->dev_err) : 0u;
    if (!e) return MPG_OK;
    *reinterpret_cast<volatile unsigned int*>(ctx->dev_err) = 0u;
    return fail(ctx, MPG_ERR_STATE, std::string("a device-side wait on a peer GPU timed out (") + ((e & DEV_ERR_REDUCE_TIMEOUT) ? "all-reduce " : "") +
                                        ((e & DEV_ERR_HALO_TIMEOUT) ? "halo " : "") + "): a rank died, diverged or issued a different launch sequence");
}
}  // namespace mpg

// =====================================================================================================
// context
// =====================================================================================================
extern "C" const char* mpg_version(void) { return "mpgmres_b200 0.1 (sm_100a)"; }

extern "C" int mpg_ctx_create(int device, mpg_ctx** out) {
    if (!out) return MPG_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0 || device >= ndev) {
        // no CPU fallback: the product path fails loudly without a CUDA device
        fprintf(stderr, "mpgmres_b200: no usable CUDA device (%s)\n", e == cudaSuccess ? "device index out of range" : cudaGetErrorString(e));
        return MPG_ERR_CUDA;
    }
    mpg_ctx* ctx = new mpg_ctx();
    ctx->device = device;
    MPG_CUDA(ctx, cudaSetDevice(device));
    cudaDeviceProp prop;
    MPG_CUDA(ctx, cudaGetDeviceProperties(&prop, device));
    ctx->num_sms = prop.multiProcessorCount;
    MPG_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
    ctx->stream = ctx->own_stream;
    MPG_CUDA(ctx, cudaMalloc(&ctx->partials, sizeof(double) * (size_t)kMaxPartBlocks * (kMaxCols + 8)));
    MPG_CUDA(ctx, cudaMalloc(&ctx->ticket, sizeof(unsigned int) * 4));
    MPG_CUDA(ctx, cudaMemset(ctx->ticket, 0, sizeof(unsigned int) * 4));
    MPG_CUDA(ctx, cudaMalloc(&ctx->dscal, sizeof(double) * 1024));
    MPG_CUDA(ctx, cudaMemset(ctx->dscal, 0, sizeof(double) * 1024));
    MPG_CUDA(ctx, cudaMallocHost(&ctx->hscal, sizeof(double) * 64));
    MPG_CUDA(ctx, cudaMalloc(&ctx->red_raw, sizeof(double) * (kMaxCols + 8)));
    MPG_CUDA(ctx, cudaHostAlloc(&ctx->dev_err, sizeof(unsigned int) * 4, cudaHostAllocMapped));
    ctx->dev_err[0] = 0;
    MPG_CUDA(ctx, cudaHostGetDevicePointer(&ctx->dev_err_d, ctx->dev_err, 0));
    {   // keep freed pool memory cached (see pool_alloc)
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess && pool) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    *out = ctx;
    return MPG_OK;
}

extern "C" int mpg_ctx_destroy(mpg_ctx* ctx) {
    if (!ctx) return MPG_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->ws && ctx->ws_free) ctx->ws_free(ctx->ws);
    if (ctx->ortho_cache && ctx->ortho_cache_free) ctx->ortho_cache_free(ctx->ortho_cache);
    for (auto& r : ctx->prof_pending) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : ctx->prof_pool) cudaEventDestroy(e);
    cudaFree(ctx->arena);
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    cudaFreeHost(ctx->stage32);
    pool_trim(ctx->device);
    cudaFree(ctx->red_raw);
    cudaFree(ctx->partials);
    cudaFree(ctx->ticket);
    cudaFree(ctx->dscal);
    cudaFreeHost(ctx->hscal);
    cudaFreeHost(ctx->dev_err);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
    return MPG_OK;
}

extern "C" int mpg_ctx_set_stream(mpg_ctx* ctx, void* s) {
    if (!ctx) return MPG_ERR_ARG;
    if (ctx->stream != (cudaStream_t)s) mpg::pool_stream_changed();   // blocks handed out so far are freed with a device-wide wait
    ctx->stream = (cudaStream_t)s;  // NULL is a valid handle: the CUDA legacy default stream
    return MPG_OK;
}
extern "C" int mpg_ctx_use_own_stream(mpg_ctx* ctx) {
    if (!ctx) return MPG_ERR_ARG;
    if (ctx->stream != ctx->own_stream) mpg::pool_stream_changed();
    ctx->stream = ctx->own_stream;
    return MPG_OK;
}
extern "C" void* mpg_ctx_stream(mpg_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" int mpg_sync(mpg_ctx* ctx) {
    if (!ctx) return MPG_ERR_ARG;
    MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return mpg::check_dev_err(ctx);
}
extern "C" const char* mpg_last_error(mpg_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }
extern "C" int mpg_num_sms(mpg_ctx* ctx) { return ctx ? ctx->num_sms : 0; }
extern "C" int64_t mpg_launch_count(mpg_ctx* ctx) { return ctx ? ctx->launches : 0; }

// one table for set / get: the name of a knob and where it lives
static int* tuning_slot(mpg_ctx* ctx, const std::string& k) {
    mpg::Tuning& t = ctx->tune;
#define MPG_KNOB(name) if (k == #name) return &t.name
    MPG_KNOB(spmv_ctas_per_sm); MPG_KNOB(dist_peer_reduce); MPG_KNOB(dist_peer_halo); MPG_KNOB(dist_overlap); MPG_KNOB(spmv_packed);
    MPG_KNOB(use_pdl); MPG_KNOB(pdl_max_rows); MPG_KNOB(fuse_tail); MPG_KNOB(vpass_stages); MPG_KNOB(fuse_min_cols);
    MPG_KNOB(vdirect_max_cols_a); MPG_KNOB(vdirect_max_cols_b); MPG_KNOB(vrow_max_cols); MPG_KNOB(vrow_max_cols_a); MPG_KNOB(gemvt_rb);
    MPG_KNOB(gemvt_rows_per_block); MPG_KNOB(passA_rb); MPG_KNOB(cgs2_fused); MPG_KNOB(vpass_serpentine); MPG_KNOB(gemvn_ctas_per_sm);
    MPG_KNOB(red_ctas_per_sm); MPG_KNOB(residual_packed); MPG_KNOB(values_static); MPG_KNOB(spmv_sigma); MPG_KNOB(sell_lpt); MPG_KNOB(mgs_fused);
    MPG_KNOB(dist_fuse_halo); MPG_KNOB(spin_limit_ms); MPG_KNOB(lookahead); MPG_KNOB(sell_variant); MPG_KNOB(sell_block); MPG_KNOB(trace);
    MPG_KNOB(dist_spmv_one_launch); MPG_KNOB(dist_push_in_spmv); MPG_KNOB(dist_ll_reduce); MPG_KNOB(host_overlap); MPG_KNOB(host_threads); MPG_KNOB(host_overlap_min_nnz);
#undef MPG_KNOB
    return nullptr;
}
extern "C" int mpg_set_tuning(mpg_ctx* ctx, const char* key, int value) {
    if (!ctx || !key) return MPG_ERR_ARG;
    const std::string k(key);
    int* slot = tuning_slot(ctx, k);
    if (!slot) return fail(ctx, MPG_ERR_ARG, "unknown tuning key " + k);
    if (k == "vrow_max_cols" || k == "vrow_max_cols_a") value = std::max(0, std::min(value, 64));
    *slot = value;
    return MPG_OK;
}
extern "C" int mpg_get_tuning(mpg_ctx* ctx, const char* key, int* value) {
    if (!ctx || !key || !value) return MPG_ERR_ARG;
    int* slot = tuning_slot(ctx, key);
    if (!slot) return fail(ctx, MPG_ERR_ARG, std::string("unknown tuning key ") + key);
    *value = *slot;
    return MPG_OK;
}

extern "C" int mpg_debug_timing(mpg_ctx* ctx, unsigned long long* device_buf) {
    if (!ctx) return MPG_ERR_ARG;
    ctx->dbg = device_buf;
    return MPG_OK;
}

extern "C" int mpg_prof_enable(mpg_ctx* ctx, int on) {
    if (!ctx) return MPG_ERR_ARG;
    ctx->prof_on = on != 0;
    return MPG_OK;
}
static int prof_fold(mpg_ctx* ctx) {
    MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto& r : ctx->prof_pending) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        ctx->prof_ms[r.cls] += ms;
        ctx->prof_bytes[r.cls] += r.bytes;
        ctx->prof_launches[r.cls] += 1;
        ctx->prof_pool.push_back(r.a);
        ctx->prof_pool.push_back(r.b);
    }
    ctx->prof_pending.clear();
    return MPG_OK;
}
extern "C" int mpg_prof_reset(mpg_ctx* ctx) {
    if (!ctx) return MPG_ERR_ARG;
    MPG_TRY(prof_fold(ctx));
    for (int c = 0; c < MPG_PROF_NCLASS; ++c) { ctx->prof_ms[c] = 0; ctx->prof_bytes[c] = 0; ctx->prof_launches[c] = 0; }
    return MPG_OK;
}
extern "C" int mpg_prof_get(mpg_ctx* ctx, int cls, double* ms, double* bytes, int64_t* launches) {
    if (!ctx || cls < 0 || cls >= MPG_PROF_NCLASS) return MPG_ERR_ARG;
    MPG_TRY(prof_fold(ctx));
    if (ms) *ms = ctx->prof_ms[cls];
    if (bytes) *bytes = ctx->prof_bytes[cls];
    if (launches) *launches = ctx->prof_launches[cls];
    return MPG_OK;
}

// =====================================================================================================
// device memory
// =====================================================================================================
extern "C" int mpg_malloc(mpg_ctx* ctx, size_t bytes, void** dptr) {
    MPG_REQUIRE(ctx, dptr != nullptr, "mpg_malloc: null out pointer");
    *dptr = nullptr;
    if (bytes == 0) return MPG_OK;
    MPG_CUDA(ctx, cudaMalloc(dptr, bytes));
    MPG_CUDA(ctx, cudaMemsetAsync(*dptr, 0, bytes, ctx->stream));
    return MPG_OK;
}
extern "C" int mpg_free(mpg_ctx* ctx, void* dptr) {
    if (!dptr) return MPG_OK;
    MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    MPG_CUDA(ctx, cudaFree(dptr));
    return MPG_OK;
}
extern "C" int mpg_memcpy_h2d(mpg_ctx* ctx, void* dst, const void* src, size_t bytes) {
    MPG_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return MPG_OK;
}
extern "C" int mpg_memcpy_d2h(mpg_ctx* ctx, void* dst, const void* src, size_t bytes) {
    MPG_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MPG_OK;
}
extern "C" int mpg_memcpy_d2d(mpg_ctx* ctx, void* dst, const void* src, size_t bytes) {
    MPG_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return MPG_OK;
}
extern "C" int mpg_memset_zero(mpg_ctx* ctx, void* dst, size_t bytes) {
    MPG_CUDA(ctx, cudaMemsetAsync(dst, 0, bytes, ctx->stream));
    return MPG_OK;
}

extern "C" int mpg_rand_vect_host(int64_t n, uint32_t seed, double* out) {
    if (!out || n < 0) return MPG_ERR_ARG;
    // gmres_perf_test.cpp:39-51: floats are drawn so x is identical whether it is later used as fp32 or fp64
    std::mt19937 engine(seed);
    std::uniform_real_distribution<float> dist;
    for (int64_t i = 0; i < n; ++i) out[i] = dist(engine);
    return MPG_OK;
}

// =====================================================================================================
// reductions: dot, nrm2
// =====================================================================================================
namespace {

constexpr int RED_THREADS = 256;

template <class T, bool IS_NRM2>
__global__ void __launch_bounds__(RED_THREADS) reduce_kernel(int64_t n, const T* __restrict__ x, const T* __restrict__ y,
                                                              double* partials, unsigned int* ticket, Epi epi) {
    constexpr int VEC = 16 / sizeof(T);
    using V = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gstride = (int64_t)gridDim.x * blockDim.x;
    const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    // dot: products accumulated in T (what BLAS ?dot does).  nrm2: squares accumulated in DOUBLE - for fp32 data that is exact
    // scaling-free protection over the whole fp32 range (squares of 1e-45 .. 3e38 neither underflow nor overflow in fp64), the
    // job BLAS ?nrm2 does with its scale / ssq recurrence (kernels_mkl.cpp:97-115); the kernel is HBM-bound either way.
    // fp64 nrm2 is the plain sum of squares: safe for |x| in [1e-150, 1e150] (stated in the header).
    using AT = typename std::conditional<IS_NRM2, double, T>::type;
    AT acc[VEC];
#pragma unroll
    for (int c = 0; c < VEC; ++c) acc[c] = AT(0);
    int64_t done = 0;
    if (aligned) {
        const int64_t nv = n / VEC;
        const V* xv = reinterpret_cast<const V*>(x);
        const V* yv = reinterpret_cast<const V*>(y);
        for (int64_t i = gtid; i < nv; i += gstride) {
            const V a = ldg_stream(xv + i);
            const V b = IS_NRM2 ? a : ldg_stream(yv + i);
            const T* pa = reinterpret_cast<const T*>(&a);
            const T* pb = reinterpret_cast<const T*>(&b);
#pragma unroll
            for (int c = 0; c < VEC; ++c) acc[c] = fma((AT)pa[c], (AT)pb[c], acc[c]);
        }
        done = nv * VEC;
    }
    for (int64_t i = done + gtid; i < n; i += gstride) {
        const T a = x[i];
        const T b = IS_NRM2 ? a : y[i];
        acc[0] = fma((AT)a, (AT)b, acc[0]);
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < VEC; ++c) s += (double)acc[c];
    s = warp_sum(s);
    __shared__ double wsum[RED_THREADS / 32];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < RED_THREADS / 32; ++w) t += wsum[w];
        partials[blockIdx.x] = t;
    }
    if (grid_last_block(ticket)) last_block_finish<T>(epi, partials, 1, gridDim.x, 1);
}

// One modified Gram-Schmidt step fused pairwise (MGS_Kernel::orthogonalize, Orthogonalization.hpp:98-106): w <- w - h_j v_j (the
// naxpy of column j) and the partial sums of h_{j+1} = v_{j+1} . w (the dot of column j + 1) in ONE pass over w: 4 n s bytes per
// column instead of 5 n s, half the launches.  Same element-to-thread mapping, the same fma per element and the same reduction
// tree as the stand-alone reduce_kernel / ew_kernel pair: results are bit-identical to the unfused sequence.
template <class T>
__global__ void __launch_bounds__(RED_THREADS) mgs_step_kernel(int64_t n, const T* __restrict__ vj, const T* __restrict__ vj1, T* w,
                                                                const T* __restrict__ hj_dev, double* partials, unsigned int* ticket, Epi epi) {
    constexpr int VEC = 16 / sizeof(T);
    using V = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
    pdl_trigger_early(n);
    pdl_wait();
    const T hj = ld_fresh(hj_dev);
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gstride = (int64_t)gridDim.x * blockDim.x;
    T acc[VEC];
#pragma unroll
    for (int c = 0; c < VEC; ++c) acc[c] = T(0);
    const int64_t nv = n / VEC;
    for (int64_t i = gtid; i < nv; i += gstride) {
        const V a = ldg_stream(reinterpret_cast<const V*>(vj) + i);
        const V b = ldg_stream(reinterpret_cast<const V*>(vj1) + i);
        V c = reinterpret_cast<V*>(w)[i];
        const T* pa = reinterpret_cast<const T*>(&a);
        const T* pb = reinterpret_cast<const T*>(&b);
        T* pc = reinterpret_cast<T*>(&c);
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
            pc[q] = fma(-hj, pa[q], pc[q]);          // naxpy(h(j,k), v_col, w)
            acc[q] = fma(pc[q], pb[q], acc[q]);      // dot(w, v_col + 1)
        }
        reinterpret_cast<V*>(w)[i] = c;
    }
    for (int64_t i = nv * VEC + gtid; i < n; i += gstride) {
        const T wn = fma(-hj, vj[i], w[i]);
        w[i] = wn;
        acc[0] = fma(wn, vj1[i], acc[0]);
    }
    double sred = 0.0;
#pragma unroll
    for (int c = 0; c < VEC; ++c) sred += (double)acc[c];
    sred = warp_sum(sred);
    __shared__ double wsum[RED_THREADS / 32];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = sred;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int q = 0; q < RED_THREADS / 32; ++q) t += wsum[q];
        partials[blockIdx.x] = t;
    }
    pdl_trigger();
    if (grid_last_block(ticket)) last_block_finish<T>(epi, partials, 1, gridDim.x, 1);
}

template <class T, bool IS_NRM2>
int launch_reduce(mpg_ctx* ctx, int64_t n, const T* x, const T* y, T* out_dev) {
    const int64_t per_block = RED_THREADS * (16 / sizeof(T)) * 4;
    int grid = (int)std::min<int64_t>(std::max<int64_t>(1, cdiv(n, per_block)), (int64_t)ctx->num_sms * ctx->tune.red_ctas_per_sm);
    grid = std::min(grid, kMaxPartBlocks);
    ProfScope prof(ctx, MPG_PROF_REDUCE, (double)n * sizeof(T) * (IS_NRM2 ? 1 : 2));
    const Epi epi = make_epi(ctx, IS_NRM2 ? EPI_NRM2 : EPI_DOT, out_dev, nullptr, 0.0, 0.0);
    reduce_kernel<T, IS_NRM2><<<grid, RED_THREADS, 0, ctx->stream>>>(n, x, y, ctx->partials, ctx->ticket, epi);
    MPG_CHECK_LAUNCH(ctx);
    return dist_finish_reduction(ctx, epi, 1, (int)sizeof(T));
}

template <class T>
int to_host(mpg_ctx* ctx, const T* dev, T* host) {
    MPG_CUDA(ctx, cudaMemcpyAsync(ctx->hscal, dev, sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(host, ctx->hscal, sizeof(T));
    return MPG_OK;
}

}  // namespace

namespace mpg {
// internal entry points used by solver.cu
int dot_dev(mpg_ctx* ctx, int64_t n, const float* x, const float* y, float* out) { return launch_reduce<float, false>(ctx, n, x, y, out); }
int dot_dev(mpg_ctx* ctx, int64_t n, const double* x, const double* y, double* out) { return launch_reduce<double, false>(ctx, n, x, y, out); }
int nrm2_dev(mpg_ctx* ctx, int64_t n, const float* x, float* out) { return launch_reduce<float, true>(ctx, n, x, x, out); }
int nrm2_dev(mpg_ctx* ctx, int64_t n, const double* x, double* out) { return launch_reduce<double, true>(ctx, n, x, x, out); }
// w -= h_j v_j ; h_{j+1} = v_{j+1} . w   (all four pointers 16-byte aligned; same grid as the stand-alone dot)
template <class T>
int mgs_step(mpg_ctx* ctx, int64_t n, const T* vj, const T* vj1, T* w, const T* hj_dev, T* hj1_dev) {
    const int64_t per_block = RED_THREADS * (16 / sizeof(T)) * 4;
    int grid = (int)std::min<int64_t>(std::max<int64_t>(1, cdiv(n, per_block)), (int64_t)ctx->num_sms * ctx->tune.red_ctas_per_sm);
    grid = std::min(grid, kMaxPartBlocks);
    ProfScope prof(ctx, MPG_PROF_REDUCE, 4.0 * (double)n * sizeof(T));
    const Epi epi = make_epi(ctx, EPI_DOT, hj1_dev, nullptr, 0.0, 0.0);
    MPG_CUDA(ctx, launch_pdl(ctx, n, mgs_step_kernel<T>, grid, RED_THREADS, 0, n, vj, vj1, w, hj_dev, ctx->partials, ctx->ticket, epi));
    MPG_CHECK_LAUNCH(ctx);
    return dist_finish_reduction(ctx, epi, 1, (int)sizeof(T));
}
template int mgs_step<float>(mpg_ctx*, int64_t, const float*, const float*, float*, const float*, float*);
template int mgs_step<double>(mpg_ctx*, int64_t, const double*, const double*, double*, const double*, double*);
}  // namespace mpg

#define MPG_DEF_RED(SFX, T)                                                                                          \
    extern "C" int mpg_dot_dev_##SFX(mpg_ctx* ctx, int64_t n, const T* x, const T* y, T* r) {                        \
        MPG_REQUIRE(ctx, n >= 0 && r, "dot: bad args");                                                              \
        return launch_reduce<T, false>(ctx, n, x, y, r);                                                             \
    }                                                                                                                \
    extern "C" int mpg_nrm2_dev_##SFX(mpg_ctx* ctx, int64_t n, const T* x, T* r) {                                   \
        MPG_REQUIRE(ctx, n >= 0 && r, "nrm2: bad args");                                                             \
        return launch_reduce<T, true>(ctx, n, x, x, r);                                                              \
    }                                                                                                                \
    extern "C" int mpg_dot_##SFX(mpg_ctx* ctx, int64_t n, const T* x, const T* y, T* r) {                            \
        MPG_REQUIRE(ctx, n >= 0 && r, "dot: bad args");                                                              \
        MPG_TRY((launch_reduce<T, false>(ctx, n, x, y, reinterpret_cast<T*>(ctx->dscal))));                          \
        return to_host<T>(ctx, reinterpret_cast<T*>(ctx->dscal), r);                                                 \
    }                                                                                                                \
    extern "C" int mpg_nrm2_##SFX(mpg_ctx* ctx, int64_t n, const T* x, T* r) {                                       \
        MPG_REQUIRE(ctx, n >= 0 && r, "nrm2: bad args");                                                             \
        MPG_TRY((launch_reduce<T, true>(ctx, n, x, x, reinterpret_cast<T*>(ctx->dscal))));                           \
        return to_host<T>(ctx, reinterpret_cast<T*>(ctx->dscal), r);                                                 \
    }
MPG_DEF_RED(f32, float)
MPG_DEF_RED(f64, double)

// =====================================================================================================
// element-wise kernels.  One template: OP selects the arithmetic; 16-byte vector path when every
// pointer is 16-byte aligned, scalar otherwise.  The scalar `alpha` comes from the host or, for the
// Scalar<T,Device> overloads, from device memory (read once per thread through the read-only path).
// =====================================================================================================
namespace {

enum EwOp { EW_AXPY, EW_NAXPY, EW_SCAL, EW_COPY, EW_FILL, EW_GDMV };

template <class TX, class TY, int OP>
__device__ __forceinline__ TY ew_apply(TX x, TY y, TY alpha, TY beta, TY d) {
    if (OP == EW_AXPY) return fma(alpha, (TY)x, y);          // y += alpha*x   (cublas?axpy)
    if (OP == EW_NAXPY) return fma(-alpha, (TY)x, y);        // y -= alpha*x   (kernels_cuda.cpp:264-288)
    if (OP == EW_SCAL) return alpha * (TY)x;                 // y = alpha*x    (copy + cublas?scal)
    if (OP == EW_COPY) return (TY)x;                         // RN conversion  (kernels.hpp:11-20)
    if (OP == EW_FILL) return alpha;
    // gdmv, kernels.hpp:143-145: beta*y + (alpha*diag)*x, every operation rounded (no contraction) so the
    // result does not depend on the compiler's fma choices
    if (sizeof(TY) == 4) return (TY)__fadd_rn(__fmul_rn((float)beta, (float)y), __fmul_rn(__fmul_rn((float)alpha, (float)d), (float)x));
    return (TY)__dadd_rn(__dmul_rn((double)beta, (double)y), __dmul_rn(__dmul_rn((double)alpha, (double)d), (double)x));
}

template <class T>
__device__ __forceinline__ void load4(const T* p, T v[4]) {
    if (sizeof(T) == 4) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = (T)t.x; v[1] = (T)t.y; v[2] = (T)t.z; v[3] = (T)t.w;
    } else {
        const double2 a = *reinterpret_cast<const double2*>(p);
        const double2 b = *reinterpret_cast<const double2*>(p + 2);
        v[0] = (T)a.x; v[1] = (T)a.y; v[2] = (T)b.x; v[3] = (T)b.y;
    }
}
template <class T>
__device__ __forceinline__ void store4(T* p, const T v[4]) {
    if (sizeof(T) == 4) {
        *reinterpret_cast<float4*>(p) = make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
    } else {
        *reinterpret_cast<double2*>(p) = make_double2((double)v[0], (double)v[1]);
        *reinterpret_cast<double2*>(p + 2) = make_double2((double)v[2], (double)v[3]);
    }
}

template <class TX, class TY, int OP>
__global__ void __launch_bounds__(256) ew_kernel(int64_t n, TY alpha, const TY* __restrict__ alpha_dev, TY beta,
                                                  const TX* x, const TY* diag, TY* y, int aligned) {
    pdl_trigger_early(n);
    pdl_wait();
    if (alpha_dev) alpha = ld_fresh(alpha_dev);
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t gstride = (int64_t)gridDim.x * blockDim.x;
    constexpr bool READS_X = (OP != EW_FILL);
    constexpr bool READS_Y = (OP == EW_AXPY || OP == EW_NAXPY || OP == EW_GDMV);
    constexpr bool READS_D = (OP == EW_GDMV);
    for (int64_t i0 = gtid * 4; i0 < n; i0 += gstride * 4) {
        TX xv[4];
        TY yv[4], dv[4];
        const int cnt = (int)min((int64_t)4, n - i0);
        if (aligned && cnt == 4) {
            if (READS_X) load4<TX>(x + i0, xv);
            if (READS_Y) load4<TY>(y + i0, yv);
            if (READS_D) load4<TY>(diag + i0, dv);
#pragma unroll
            for (int c = 0; c < 4; ++c) yv[c] = ew_apply<TX, TY, OP>(xv[c], yv[c], alpha, beta, dv[c]);
            store4<TY>(y + i0, yv);
        } else {
            for (int c = 0; c < cnt; ++c) {
                const TX xs = READS_X ? x[i0 + c] : TX(0);
                const TY ys = READS_Y ? y[i0 + c] : TY(0);
                const TY ds = READS_D ? diag[i0 + c] : TY(0);
                y[i0 + c] = ew_apply<TX, TY, OP>(xs, ys, alpha, beta, ds);
            }
        }
    }
}

template <class TX, class TY, int OP>
int launch_ew(mpg_ctx* ctx, int64_t n, TY alpha, const TY* alpha_dev, TY beta, const TX* x, const TY* diag, TY* y) {
    if (n <= 0) return MPG_OK;
    const int grid = (int)std::min<int64_t>(cdiv(n, 256 * 4), (int64_t)ctx->num_sms * 16);
    const int aligned = (((uintptr_t)x | (uintptr_t)diag | (uintptr_t)y) & 15) == 0;
    const double ew_bytes = (double)n * ((OP != EW_FILL ? sizeof(TX) : 0) + sizeof(TY) * (1 + (OP == EW_AXPY || OP == EW_NAXPY || OP == EW_GDMV) + (OP == EW_GDMV)));
    ProfScope prof(ctx, MPG_PROF_ELEMENTWISE, ew_bytes);
    MPG_CUDA(ctx, launch_pdl(ctx, n, ew_kernel<TX, TY, OP>, grid, 256, 0, n, alpha, alpha_dev, beta, x, diag, y, aligned));
    MPG_CHECK_LAUNCH(ctx);
    return MPG_OK;
}

}  // namespace

namespace mpg {
int scal_host(mpg_ctx* ctx, int64_t n, float a, const float* x, float* y) { return launch_ew<float, float, EW_SCAL>(ctx, n, a, nullptr, 0.f, x, nullptr, y); }
int scal_host(mpg_ctx* ctx, int64_t n, double a, const double* x, double* y) { return launch_ew<double, double, EW_SCAL>(ctx, n, a, nullptr, 0.0, x, nullptr, y); }
int scal_devp(mpg_ctx* ctx, int64_t n, const float* a, const float* x, float* y) { return launch_ew<float, float, EW_SCAL>(ctx, n, 0.f, a, 0.f, x, nullptr, y); }
int scal_devp(mpg_ctx* ctx, int64_t n, const double* a, const double* x, double* y) { return launch_ew<double, double, EW_SCAL>(ctx, n, 0.0, a, 0.0, x, nullptr, y); }
int naxpy_devp(mpg_ctx* ctx, int64_t n, const float* a, const float* x, float* y) { return launch_ew<float, float, EW_NAXPY>(ctx, n, 0.f, a, 0.f, x, nullptr, y); }
int naxpy_devp(mpg_ctx* ctx, int64_t n, const double* a, const double* x, double* y) { return launch_ew<double, double, EW_NAXPY>(ctx, n, 0.0, a, 0.0, x, nullptr, y); }
int axpy_host(mpg_ctx* ctx, int64_t n, float a, const float* x, float* y) { return launch_ew<float, float, EW_AXPY>(ctx, n, a, nullptr, 0.f, x, nullptr, y); }
int axpy_host(mpg_ctx* ctx, int64_t n, double a, const double* x, double* y) { return launch_ew<double, double, EW_AXPY>(ctx, n, a, nullptr, 0.0, x, nullptr, y); }
int cast_copy(mpg_ctx* ctx, int64_t n, const double* x, float* y) { return launch_ew<double, float, EW_COPY>(ctx, n, 0.f, nullptr, 0.f, x, nullptr, y); }
int cast_copy(mpg_ctx* ctx, int64_t n, const float* x, double* y) { return launch_ew<float, double, EW_COPY>(ctx, n, 0.0, nullptr, 0.0, x, nullptr, y); }
int cast_copy(mpg_ctx* ctx, int64_t n, const float* x, float* y) { return launch_ew<float, float, EW_COPY>(ctx, n, 0.f, nullptr, 0.f, x, nullptr, y); }
int cast_copy(mpg_ctx* ctx, int64_t n, const double* x, double* y) { return launch_ew<double, double, EW_COPY>(ctx, n, 0.0, nullptr, 0.0, x, nullptr, y); }
int fill_host(mpg_ctx* ctx, int64_t n, float a, float* x) { return launch_ew<float, float, EW_FILL>(ctx, n, a, nullptr, 0.f, nullptr, nullptr, x); }
int fill_host(mpg_ctx* ctx, int64_t n, double a, double* x) { return launch_ew<double, double, EW_FILL>(ctx, n, a, nullptr, 0.0, nullptr, nullptr, x); }
int gdmv_host(mpg_ctx* ctx, int64_t n, float a, const float* d, const float* x, float b, float* y) { return launch_ew<float, float, EW_GDMV>(ctx, n, a, nullptr, b, x, d, y); }
int gdmv_host(mpg_ctx* ctx, int64_t n, double a, const double* d, const double* x, double b, double* y) { return launch_ew<double, double, EW_GDMV>(ctx, n, a, nullptr, b, x, d, y); }
}  // namespace mpg

#define MPG_DEF_EW(SFX, T)                                                                                              \
    extern "C" int mpg_axpy_##SFX(mpg_ctx* c, int64_t n, T a, const T* x, T* y) { return mpg::axpy_host(c, n, a, x, y); } \
    extern "C" int mpg_axpy_dev_##SFX(mpg_ctx* c, int64_t n, const T* a, const T* x, T* y) {                            \
        return launch_ew<T, T, EW_AXPY>(c, n, T(0), a, T(0), x, nullptr, y);                                            \
    }                                                                                                                   \
    extern "C" int mpg_naxpy_dev_##SFX(mpg_ctx* c, int64_t n, const T* a, const T* x, T* y) { return mpg::naxpy_devp(c, n, a, x, y); } \
    extern "C" int mpg_scal_##SFX(mpg_ctx* c, int64_t n, T a, const T* x, T* y) { return mpg::scal_host(c, n, a, x, y); } \
    extern "C" int mpg_scal_dev_##SFX(mpg_ctx* c, int64_t n, const T* a, const T* x, T* y) { return mpg::scal_devp(c, n, a, x, y); } \
    extern "C" int mpg_fill_##SFX(mpg_ctx* c, int64_t n, T a, T* x) { return mpg::fill_host(c, n, a, x); }              \
    extern "C" int mpg_gdmv_##SFX(mpg_ctx* c, int64_t n, T a, const T* d, const T* x, T b, T* y) { return mpg::gdmv_host(c, n, a, d, x, b, y); }
MPG_DEF_EW(f32, float)
MPG_DEF_EW(f64, double)

extern "C" int mpg_copy_f32_f32(mpg_ctx* c, int64_t n, const float* x, float* y) { return mpg::cast_copy(c, n, x, y); }
extern "C" int mpg_copy_f64_f64(mpg_ctx* c, int64_t n, const double* x, double* y) { return mpg::cast_copy(c, n, x, y); }
extern "C" int mpg_copy_f64_f32(mpg_ctx* c, int64_t n, const double* x, float* y) { return mpg::cast_copy(c, n, x, y); }
extern "C" int mpg_copy_f32_f64(mpg_ctx* c, int64_t n, const float* x, double* y) { return mpg::cast_copy(c, n, x, y); }

// =====================================================================================================
// Givens rotations, triangular solve (tiny, latency-bound; one warp or one thread each)
// =====================================================================================================
namespace {

// netlib ?rotg followed by b := 0 (kernels_cuda.cpp:394-420).  No FMA contraction: written with
// explicit intrinsics so the rounding sequence equals the oracle's (a/scale, squares, sum, sqrt, product).
template <class T> __device__ __forceinline__ T mul_rn(T a, T b);
template <> __device__ __forceinline__ float mul_rn<float>(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double mul_rn<double>(double a, double b) { return __dmul_rn(a, b); }
template <class T> __device__ __forceinline__ T add_rn(T a, T b);
template <> __device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }
template <class T> __device__ __forceinline__ T div_rn(T a, T b);
template <> __device__ __forceinline__ float div_rn<float>(float a, float b) { return __fdiv_rn(a, b); }
template <> __device__ __forceinline__ double div_rn<double>(double a, double b) { return __ddiv_rn(a, b); }
template <class T> __device__ __forceinline__ T sqrt_rn(T a);
template <> __device__ __forceinline__ float sqrt_rn<float>(float a) { return __fsqrt_rn(a); }
template <> __device__ __forceinline__ double sqrt_rn<double>(double a) { return __dsqrt_rn(a); }

template <class T>
__device__ __forceinline__ void dev_rotg(T& a, T& b, T& c, T& s) {
    const T roe = (fabs(a) > fabs(b)) ? a : b;
    const T scale = add_rn(fabs(a), fabs(b));
    T r;
    if (scale == T(0)) {
        c = T(1); s = T(0); r = T(0);
    } else {
        const T as = div_rn(a, scale), bs = div_rn(b, scale);
        r = mul_rn(scale, sqrt_rn(add_rn(mul_rn(as, as), mul_rn(bs, bs))));
        r = mul_rn(copysign(T(1), roe), r);
        c = div_rn(a, r);
        s = div_rn(b, r);
    }
    a = r;
    b = T(0);
}
// ?rot with n = 1: (a,b) <- (c a + s b, c b - s a), products and sums rounded separately like the oracle
template <class T>
__device__ __forceinline__ void dev_rot(T& a, T& b, T c, T s) {
    const T t = add_rn(mul_rn(c, a), mul_rn(s, b));
    b = add_rn(mul_rn(c, b), -mul_rn(s, a));
    a = t;
}

template <class T>
__global__ void rotg_kernel(T* a, T* b, T* c, T* s) {
    T va = *a, vb = *b, vc, vs;
    dev_rotg(va, vb, vc, vs);
    *a = va; *b = vb; *c = vc; *s = vs;
}
template <class T>
__global__ void rot_kernel(T* a, T* b, const T* c, const T* s) {
    T va = *a, vb = *b;
    dev_rot(va, vb, *c, *s);
    *a = va; *b = vb;
}
template <class T>
__global__ void rot_vec_kernel(int64_t k, T* a, const T* c, const T* s) {
    if (k <= 0) return;
    T cur = a[0];
    for (int64_t j = 0; j < k; ++j) {
        T nxt = a[j + 1];
        dev_rot(cur, nxt, c[j], s[j]);
        a[j] = cur;
        cur = nxt;
    }
    a[k] = cur;
}
// gmres.cpp:219-226 in one launch.  One warp: lanes prefetch c/s/h into shared memory, lane 0 runs the
// dependent rotation chain from shared memory.
template <class T>
__device__ __forceinline__ void givens_step_body(unsigned char* smem_raw, int64_t k, T* h, int64_t ldh, T* cs, T* sn, T* s, double* resid,
                                                 double* resid_host) {
    T* sh = reinterpret_cast<T*>(smem_raw);        // k+2
    T* sc = sh + (k + 2);                          // k
    T* ss = sc + k;                                // k
    T* hcol = h + k * ldh;
    for (int64_t j = threadIdx.x; j < k + 2; j += 32) sh[j] = ld_fresh(hcol + j);
    for (int64_t j = threadIdx.x; j < k; j += 32) { sc[j] = ld_fresh(cs + j); ss[j] = ld_fresh(sn + j); }
    __syncwarp();
    if (threadIdx.x == 0) {
        T cur = sh[0];
        for (int64_t j = 0; j < k; ++j) {
            T nxt = sh[j + 1];
            dev_rot(cur, nxt, sc[j], ss[j]);
            sh[j] = cur;
            cur = nxt;
        }
        T hk1 = sh[k + 1], c, sv;
        dev_rotg(cur, hk1, c, sv);
        sh[k] = cur;
        sh[k + 1] = hk1;
        cs[k] = c;
        sn[k] = sv;
        T s0 = s[k], s1 = s[k + 1];
        dev_rot(s0, s1, c, sv);
        s[k] = s0;
        s[k + 1] = s1;
        if (resid) *resid = fabs((double)s1);
        if (resid_host) {   // mapped pinned memory: lets the host follow the residual without synchronising the stream
            *reinterpret_cast<volatile double*>(resid_host) = fabs((double)s1);
            __threadfence_system();
        }
    }
    __syncwarp();
    for (int64_t j = threadIdx.x; j < k + 2; j += 32) hcol[j] = sh[j];
}

template <class T>
__global__ void __launch_bounds__(32) givens_step_kernel(int64_t k, T* h, int64_t ldh, T* cs, T* sn, T* s, double* resid, double* resid_host) {
    extern __shared__ unsigned char smem_raw[];
    pdl_trigger();
    pdl_wait();
    givens_step_body<T>(smem_raw, k, h, ldh, cs, sn, s, resid, resid_host);
}

// Tail of an Arnoldi step in ONE launch: the Givens update of column k (gmres.cpp:219-226) in warp 0 of block 0, the halo push of
// the new column (multi-GPU) in the next npush blocks, V(:,k+1) = w * (1/h(k+1,k))  (Orthogonalization.hpp:58-59) in all the
// others.  The three are independent: they only consume what the orthogonalisation left behind (w, 1/h(k+1,k), h(:,k)).
template <class T>
__global__ void __launch_bounds__(256) arnoldi_tail_kernel(int64_t n, const T* __restrict__ inv_dev, const T* x, T* y, int aligned, int64_t k, T* h,
                                                            int64_t ldh, T* cs, T* sn, T* s, double* resid, double* resid_host,
                                                            const __grid_constant__ PushArgs push, int npush) {
    extern __shared__ unsigned char smem_raw[];
    pdl_trigger_early(n);
    pdl_wait();
    // block 0: Givens chain; blocks 1..npush: halo push; the rest: normalisation.  The latency-bound parts (a dependent rotation chain,
    // remote stores + a system-scope fence + the flag) come FIRST in the grid so that they run under the bandwidth-bound normalisation
    // instead of after its last wave.
    if (blockIdx.x == 0) {
        if (threadIdx.x < 32) givens_step_body<T>(smem_raw, k, h, ldh, cs, sn, s, resid, resid_host);
        return;
    }
    if ((int)blockIdx.x <= npush) {
        // multi-GPU: the boundary rows of the NEW basis vector, w[idx] * (1/h), go straight into the halo tail of the neighbours'
        // copy of that column (same product as the local store below: bit-identical values on both sides)
        halo_push_block<T>(push, (int)blockIdx.x - 1, x, inv_dev);
        return;
    }
    const int new_blocks = (int)gridDim.x - 1 - npush;
    const int bid = (int)blockIdx.x - 1 - npush;
    const T alpha = ld_fresh(inv_dev);
    const int64_t gtid = (int64_t)bid * blockDim.x + threadIdx.x;
    const int64_t gstride = (int64_t)new_blocks * blockDim.x;
    int64_t i0 = gtid * 4;
    // four 16-byte loads in flight per thread (a single-wave grid, launcher below): the kernel is latency-bound at slab sizes
    if (aligned) {
        for (; i0 + 3 * gstride * 4 + 4 <= n; i0 += 4 * gstride * 4) {
            T v[4][4];
#pragma unroll
            for (int u = 0; u < 4; ++u) load4(x + i0 + (int64_t)u * gstride * 4, v[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int c = 0; c < 4; ++c) v[u][c] = alpha * v[u][c];
                store4(y + i0 + (int64_t)u * gstride * 4, v[u]);
            }
        }
    }
    for (; i0 < n; i0 += gstride * 4) {
        const int cnt = (int)min((int64_t)4, n - i0);
        if (aligned && cnt == 4) {
            T v[4];
            load4(x + i0, v);
#pragma unroll
            for (int c = 0; c < 4; ++c) v[c] = alpha * v[c];
            store4(y + i0, v);
        } else {
            for (int c = 0; c < cnt; ++c) y[i0 + c] = alpha * x[i0 + c];
        }
    }
}

// netlib ?trsv, Upper/NoTrans/NonUnit in column form (the oracle's order); Lower and Trans forms are the
// row-oriented textbook loops.  n <= restart length, one thread: the solve is a dependent chain anyway.
template <class T>
__global__ void trsv_kernel(int upper, int trans, int64_t n, const T* A, int64_t ld, T* x) {
    if (upper && !trans) {
        for (int64_t j = n; j-- > 0;) {
            if (x[j] != T(0)) {
                x[j] = div_rn(x[j], A[j + j * ld]);
                const T t = x[j];
                for (int64_t i = j; i-- > 0;) x[i] = fma(-t, A[i + j * ld], x[i]);
            }
        }
    } else if (!upper && !trans) {
        for (int64_t j = 0; j < n; ++j) {
            if (x[j] != T(0)) {
                x[j] = div_rn(x[j], A[j + j * ld]);
                const T t = x[j];
                for (int64_t i = j + 1; i < n; ++i) x[i] = fma(-t, A[i + j * ld], x[i]);
            }
        }
    } else if (upper && trans) {  // solve U^T x = b: forward substitution on columns of U
        for (int64_t j = 0; j < n; ++j) {
            T t = x[j];
            for (int64_t i = 0; i < j; ++i) t = fma(-A[i + j * ld], x[i], t);
            x[j] = div_rn(t, A[j + j * ld]);
        }
    } else {  // L^T x = b
        for (int64_t j = n; j-- > 0;) {
            T t = x[j];
            for (int64_t i = n - 1; i > j; --i) t = fma(-A[i + j * ld], x[i], t);
            x[j] = div_rn(t, A[j + j * ld]);
        }
    }
}

// Upper / NoTrans / NonUnit (the solution_update case, gmres.cpp:288,300) with the whole triangle staged in shared
// memory: same column sweep and the same per-element operation order as the netlib loop above (bit-identical
// results), but the i-loop of each column runs across the block and the matrix is read with coalesced loads.
template <class T>
__global__ void __launch_bounds__(128) trsv_upper_smem_kernel(int n, const T* __restrict__ A, int64_t ld, T* x) {
    extern __shared__ __align__(16) unsigned char smem_trsv[];
    T* As = reinterpret_cast<T*>(smem_trsv);  // n x n, column-major, ld = n
    T* xs = As + (size_t)n * n;
    for (int idx = threadIdx.x; idx < n * n; idx += blockDim.x) {
        const int i = idx % n, j = idx / n;
        As[idx] = (i <= j) ? A[i + (int64_t)j * ld] : T(0);
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) xs[i] = x[i];
    __syncthreads();
    for (int j = n - 1; j >= 0; --j) {
        if (xs[j] != T(0)) {   // uniform branch: every thread reads the same shared value
            __syncthreads();
            if (threadIdx.x == 0) xs[j] = div_rn(xs[j], As[j + j * n]);
            __syncthreads();
            const T t = xs[j];
            for (int i = threadIdx.x; i < j; i += blockDim.x) xs[i] = fma(-t, As[i + j * n], xs[i]);
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) x[i] = xs[i];
}

}  // namespace

namespace mpg {
template <class T>
int givens_step(mpg_ctx* ctx, int64_t k, T* h, int64_t ldh, T* cs, T* sn, T* s, double* resid, double* resid_host) {
    const size_t smem = sizeof(T) * (size_t)(3 * k + 2);
    ProfScope prof(ctx, MPG_PROF_SMALL, 0.0);
    MPG_CUDA(ctx, launch_pdl(ctx, (int64_t)0, givens_step_kernel<T>, 1, 32, smem, k, h, ldh, cs, sn, s, resid, resid_host));
    MPG_CHECK_LAUNCH(ctx);
    return MPG_OK;
}
template <class T>
int arnoldi_tail(mpg_ctx* ctx, int64_t n, const T* inv_dev, const T* w, T* vnext, int64_t k, T* h, int64_t ldh, T* cs, T* sn, T* s, double* resid,
                 double* resid_host, const PushArgs* push) {
    const size_t smem = sizeof(T) * (size_t)(3 * k + 2);
    PushArgs pa;
    if (push) pa = *push;
    const int npush = pa.npeers * pa.bpp;
    // normalisation blocks: one 16-byte group per thread for small operands (as many CTAs as there is work), at most half a wave
    // (4 CTAs of 256 threads per SM) with up to four groups in flight per thread for large ones
    const int grid = (int)std::min<int64_t>(std::max<int64_t>(cdiv(n, 256 * 4), 1), (int64_t)ctx->num_sms * 4) + npush + 1;
    const int aligned = (((uintptr_t)w | (uintptr_t)vnext) & 15) == 0;
    ProfScope prof(ctx, MPG_PROF_ELEMENTWISE, 2.0 * (double)n * sizeof(T));
    MPG_CUDA(ctx, launch_pdl(ctx, n, arnoldi_tail_kernel<T>, grid, 256, smem, n, inv_dev, w, vnext, aligned, k, h, ldh, cs, sn, s, resid, resid_host, pa, npush));
    MPG_CHECK_LAUNCH(ctx);
    return MPG_OK;
}
template int arnoldi_tail<float>(mpg_ctx*, int64_t, const float*, const float*, float*, int64_t, float*, int64_t, float*, float*, float*, double*, double*, const PushArgs*);
template int arnoldi_tail<double>(mpg_ctx*, int64_t, const double*, const double*, double*, int64_t, double*, int64_t, double*, double*, double*, double*, double*, const PushArgs*);
template int givens_step<float>(mpg_ctx*, int64_t, float*, int64_t, float*, float*, float*, double*, double*);
template int givens_step<double>(mpg_ctx*, int64_t, double*, int64_t, double*, double*, double*, double*, double*);
template <class T>
int trsv(mpg_ctx* ctx, int upper, int trans, int64_t n, const T* A, int64_t ld, T* x) {
    if (n <= 0) return MPG_OK;
    ProfScope prof(ctx, MPG_PROF_SMALL, 0.0);
    const size_t smem = sizeof(T) * ((size_t)n * n + n);
    if (upper && !trans && smem <= 200 * 1024) {
        auto kern = trsv_upper_smem_kernel<T>;
        if (smem > 48 * 1024) MPG_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<1, 128, smem, ctx->stream>>>((int)n, A, ld, x);
        MPG_CHECK_LAUNCH(ctx);
        return MPG_OK;
    }
    trsv_kernel<T><<<1, 1, 0, ctx->stream>>>(upper, trans, n, A, ld, x);
    MPG_CHECK_LAUNCH(ctx);
    return MPG_OK;
}
template int trsv<float>(mpg_ctx*, int, int, int64_t, const float*, int64_t, float*);
template int trsv<double>(mpg_ctx*, int, int, int64_t, const double*, int64_t, double*);
}  // namespace mpg

#define MPG_DEF_LS(SFX, T)                                                                                     \
    extern "C" int mpg_rotg_##SFX(mpg_ctx* ctx, T* a, T* b, T* c, T* s) {                                      \
        rotg_kernel<T><<<1, 1, 0, ctx->stream>>>(a, b, c, s);                                                  \
        MPG_CHECK_LAUNCH(ctx);                                                                                 \
        return MPG_OK;                                                                                         \
    }                                                                                                          \
    extern "C" int mpg_rot_##SFX(mpg_ctx* ctx, T* a, T* b, const T* c, const T* s) {                           \
        rot_kernel<T><<<1, 1, 0, ctx->stream>>>(a, b, c, s);                                                   \
        MPG_CHECK_LAUNCH(ctx);                                                                                 \
        return MPG_OK;                                                                                         \
    }                                                                                                          \
    extern "C" int mpg_rot_vec_##SFX(mpg_ctx* ctx, int64_t k, T* a, const T* c, const T* s) {                  \
        rot_vec_kernel<T><<<1, 1, 0, ctx->stream>>>(k, a, c, s);                                               \
        MPG_CHECK_LAUNCH(ctx);                                                                                 \
        return MPG_OK;                                                                                         \
    }                                                                                                          \
    extern "C" int mpg_trsv_##SFX(mpg_ctx* ctx, int upper, int trans, int64_t n, const T* A, int64_t ld, T* x) { \
        MPG_REQUIRE(ctx, n >= 0 && ld >= n, "trsv: bad dims");                                                 \
        return mpg::trsv<T>(ctx, upper, trans, n, A, ld, x);                                                   \
    }                                                                                                          \
    extern "C" int mpg_givens_step_##SFX(mpg_ctx* ctx, int64_t k, T* h, int64_t ldh, T* cs, T* sn, T* s, double* r) { \
        MPG_REQUIRE(ctx, k >= 0 && k + 2 <= kMaxCols + 2, "givens_step: bad k");                               \
        return mpg::givens_step<T>(ctx, k, h, ldh, cs, sn, s, r, nullptr);                                              \
    }
MPG_DEF_LS(f32, float)
MPG_DEF_LS(f64, double)
