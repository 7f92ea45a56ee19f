// hostpath.cu — mpg_gmres_solve_host: HOST CSR + b in, x out (the end-to-end call: H2D, plan, solve, D2H inside).
//
// Two shapes:
//   * serial: one H2D of everything, then plan, cast, solve (every mode);
//   * overlapped (mixed precision, knob host_overlap): the link is the bottleneck (5.7 GB at ~55 GB/s = 105 ms for the 16.7 M-row
//     system, before a 220 ms solve), and the first restart cycle of GMRES-IR only multiplies with the fp32 operator
//     (gmres.cpp:189-232).  So
//       - the indices go first, and while they travel the host's cores cast the fp64 values to fp32 into a pinned staging buffer
//         (the same round-to-nearest conversion SparseMatrix<float>(A) applies on the device, types_cuda.hpp:82-101);
//       - a feeder thread puts each finished fp32 chunk on the wire, filling the gaps with fp64 chunks so the link never idles;
//       - the solve starts as soon as the fp32 operator is complete; the rest of the fp64 values land during the first cycle and
//         the solver waits for them at its first fp64 residual that needs the operator (DeferredV64, solver.cu).  With x0 = 0
//         (checked here, on the host) the very first residual is r = b - A*0 = b and needs no operator.
//     Results are bit-identical to the serial shape (tests/test_solver_gpu.py).
#include <sched.h>

#include <atomic>
#include <memory>
#include <thread>
#include <vector>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "common.cuh"

using namespace mpg;

namespace mpg {
int cast_copy(mpg_ctx*, int64_t, const double*, float*);
int sell_plan_get(mpg_ctx*, const mpg_csr*, const mpg_sell_plan**);
}

namespace {

struct Operands {
    int* row_map; int* inds; double* vals; float* vals32; double* b; double* x;
};

// carve the operands out of the context's grow-only arena: repeated solves pay no allocation cost
int carve(mpg_ctx* ctx, int nrows, int64_t nnz, Operands* o) {
    auto up = [](size_t v) { return (v + 255) & ~size_t(255); };
    const size_t o_rm = 0;
    const size_t o_in = o_rm + up(sizeof(int) * (size_t)(nrows + 1));
    const size_t o_v64 = o_in + up(sizeof(int) * (size_t)nnz);
    const size_t o_v32 = o_v64 + up(sizeof(double) * (size_t)nnz);
    const size_t o_b = o_v32 + up(sizeof(float) * (size_t)nnz);
    const size_t o_x = o_b + up(sizeof(double) * (size_t)nrows);
    const size_t total = o_x + up(sizeof(double) * (size_t)nrows) + 256;
    if (total > ctx->arena_bytes) {
        MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->copy_stream) MPG_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
        cudaFree(ctx->arena);
        ctx->arena = nullptr; ctx->arena_bytes = 0;
        MPG_CUDA(ctx, cudaMalloc(&ctx->arena, total));
        ctx->arena_bytes = total;
    }
    char* base = static_cast<char*>(ctx->arena);
    o->row_map = reinterpret_cast<int*>(base + o_rm);
    o->inds = reinterpret_cast<int*>(base + o_in);
    o->vals = reinterpret_cast<double*>(base + o_v64);
    o->vals32 = reinterpret_cast<float*>(base + o_v32);
    o->b = reinterpret_cast<double*>(base + o_b);
    o->x = reinterpret_cast<double*>(base + o_x);
    return MPG_OK;
}

int host_cpus() {
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0) {
        const int c = CPU_COUNT(&set);
        if (c > 0) return c;
    }
    const unsigned h = std::thread::hardware_concurrency();
    return h ? (int)h : 1;
}

// d[i] = (float) s[i], round to nearest even (the default MXCSR mode) - the conversion cvt.rn.f32.f64 does on the device.  d is
// 16-byte aligned pinned memory that only the DMA engine reads next: streaming stores, no read-for-ownership.
void cast_range(const double* s, float* d, size_t n) {
    size_t i = 0;
#if defined(__SSE2__)
    if ((reinterpret_cast<uintptr_t>(d) & 15) == 0) {
        for (; i + 4 <= n; i += 4) {
            const __m128 lo = _mm_cvtpd_ps(_mm_loadu_pd(s + i));
            const __m128 hi = _mm_cvtpd_ps(_mm_loadu_pd(s + i + 2));
            _mm_stream_ps(d + i, _mm_movelh_ps(lo, hi));
        }
        _mm_sfence();
    }
#endif
    for (; i < n; ++i) d[i] = (float)s[i];
}

// every bit zero (+0.0 everywhere)?
bool all_plus_zero(const double* x, size_t n) {
    const uint64_t* u = reinterpret_cast<const uint64_t*>(x);
    uint64_t acc = 0;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        acc |= u[i] | u[i + 1] | u[i + 2] | u[i + 3] | u[i + 4] | u[i + 5] | u[i + 6] | u[i + 7];
        if (acc) return false;
    }
    for (; i < n; ++i) acc |= u[i];
    return acc == 0;
}

struct Events {
    std::vector<cudaEvent_t> ev;
    cudaEvent_t make() { cudaEvent_t e = nullptr; cudaEventCreate(&e); ev.push_back(e); return e; }
    ~Events() { for (auto e : ev) if (e) cudaEventDestroy(e); }
};

#define MPG_TRY_H(expr) do { rc = (expr); if (rc != MPG_OK) { cleanup(); return rc; } } while (0)
#define MPG_CUDA_H(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { (void)cudaGetLastError(); cleanup(); return fail(ctx, MPG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); } } while (0)

int solve_host_serial(mpg_ctx* ctx, const mpg_gmres_params* p, int nrows, int64_t nnz, const int* row_map_h, const int* inds_h, const double* vals64_h,
                      const double* b_h, double* x_h, mpg_gmres_stats* st, double* hist_inner, int64_t cap_inner, double* hist_outer, int64_t cap_outer) {
    Operands o;
    MPG_TRY(carve(ctx, nrows, nnz, &o));
    mpg_csr* A = nullptr;
    Events evs;
    cudaEvent_t e0 = evs.make(), e1 = evs.make(), e2 = evs.make(), e3 = evs.make();
    int rc = MPG_OK;
    auto cleanup = [&]() {
        cudaStreamSynchronize(ctx->stream);
        if (A) mpg_csr_destroy(A);
    };
    MPG_CUDA_H(cudaEventRecord(e0, ctx->stream));
    MPG_CUDA_H(cudaMemcpyAsync(o.row_map, row_map_h, sizeof(int) * (size_t)(nrows + 1), cudaMemcpyHostToDevice, ctx->stream));
    MPG_CUDA_H(cudaMemcpyAsync(o.inds, inds_h, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream));
    MPG_CUDA_H(cudaMemcpyAsync(o.vals, vals64_h, sizeof(double) * (size_t)nnz, cudaMemcpyHostToDevice, ctx->stream));
    MPG_CUDA_H(cudaMemcpyAsync(o.b, b_h, sizeof(double) * (size_t)nrows, cudaMemcpyHostToDevice, ctx->stream));
    MPG_CUDA_H(cudaMemcpyAsync(o.x, x_h, sizeof(double) * (size_t)nrows, cudaMemcpyHostToDevice, ctx->stream));
    MPG_CUDA_H(cudaEventRecord(e1, ctx->stream));
    MPG_TRY_H(mpg_csr_create(ctx, nrows, nrows, nnz, o.row_map, o.inds, &A));
    MPG_TRY_H(cast_copy(ctx, nnz, o.vals, o.vals32));   // SparseMatrix<float>(A), gmres_perf_test.cpp:136
    MPG_TRY_H(mpg_gmres_solve(ctx, p, A, o.vals, o.vals32, o.b, o.x, st, hist_inner, cap_inner, hist_outer, cap_outer));
    MPG_CUDA_H(cudaEventRecord(e2, ctx->stream));
    MPG_CUDA_H(cudaMemcpyAsync(x_h, o.x, sizeof(double) * (size_t)nrows, cudaMemcpyDeviceToHost, ctx->stream));
    MPG_CUDA_H(cudaEventRecord(e3, ctx->stream));
    MPG_CUDA_H(cudaEventSynchronize(e3));
    float t = 0.f;
    cudaEventElapsedTime(&t, e0, e1); st->h2d_ms = t; st->h2d_all_ms = t;
    cudaEventElapsedTime(&t, e2, e3); st->d2h_ms = t;
    st->h2d_bytes = (int64_t)(sizeof(int) * (size_t)(nrows + 1) + sizeof(int) * (size_t)nnz + sizeof(double) * (size_t)nnz + 2 * sizeof(double) * (size_t)nrows);
    st->host_overlap = 0;
    cleanup();
    return MPG_OK;
}

constexpr int kNeedSerial = -1000;   // internal: the overlapped shape could not get its resources, take the serial one

int solve_host_overlapped(mpg_ctx* ctx, const mpg_gmres_params* p, int nrows, int64_t nnz, const int* row_map_h, const int* inds_h,
                          const double* vals64_h, const double* b_h, double* x_h, mpg_gmres_stats* st, double* hist_inner, int64_t cap_inner,
                          double* hist_outer, int64_t cap_outer) {
    Operands o;
    MPG_TRY(carve(ctx, nrows, nnz, &o));
    if (!ctx->copy_stream && cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
        (void)cudaGetLastError();
        ctx->copy_stream = nullptr;
        return kNeedSerial;
    }
    if (ctx->stage32_elems < (size_t)nnz) {
        cudaStreamSynchronize(ctx->copy_stream);
        cudaFreeHost(ctx->stage32);
        ctx->stage32 = nullptr; ctx->stage32_elems = 0;
        if (cudaHostAlloc(&ctx->stage32, sizeof(float) * (size_t)nnz, cudaHostAllocDefault) != cudaSuccess) {
            (void)cudaGetLastError();
            ctx->stage32 = nullptr;
            return kNeedSerial;   // not enough pinnable host memory
        }
        ctx->stage32_elems = (size_t)nnz;
    }
    cudaStream_t cs = ctx->copy_stream;
    float* stage = ctx->stage32;

    constexpr int64_t CH = int64_t(1) << 21;   // values per chunk: 8 MB of fp32 (0.15 ms on the wire), 16 MB of fp64
    const int64_t nch = cdiv(nnz, CH);
    std::unique_ptr<std::atomic<unsigned char>[]> done(new std::atomic<unsigned char>[(size_t)nch]);
    for (int64_t c = 0; c < nch; ++c) done[(size_t)c].store(0, std::memory_order_relaxed);
    std::atomic<int64_t> next{0};
    std::atomic<int> v32_recorded{0};
    std::atomic<bool> abandon{false};
    DeferredV64 df;

    Events evs;
    cudaEvent_t e0 = evs.make(), e_inds = evs.make(), e_v32 = evs.make(), e2 = evs.make(), e3 = evs.make();
    df.ev = evs.make();
    cudaEvent_t ring[2] = {evs.make(), evs.make()};
    for (auto e : evs.ev) if (!e) return kNeedSerial;

    mpg_csr* A = nullptr;
    std::vector<std::thread> workers;
    std::thread feeder;
    int rc = MPG_OK;
    auto join_all = [&]() {
        for (auto& t : workers) if (t.joinable()) t.join();
        if (feeder.joinable()) feeder.join();
    };
    auto cleanup = [&]() {   // error paths abandon unclaimed chunks; after a complete solve there is nothing left to abandon
        abandon.store(true, std::memory_order_relaxed);
        join_all();
        ctx->defer = nullptr;
        cudaStreamSynchronize(cs);
        cudaStreamSynchronize(ctx->stream);
        if (A) mpg_csr_destroy(A);
    };

    // the arena may still be read by earlier work of this context: the copy stream starts behind it
    MPG_CUDA_H(cudaEventRecord(e0, ctx->stream));
    MPG_CUDA_H(cudaStreamWaitEvent(cs, e0, 0));
    MPG_CUDA_H(cudaMemcpyAsync(o.row_map, row_map_h, sizeof(int) * (size_t)(nrows + 1), cudaMemcpyHostToDevice, cs));
    MPG_CUDA_H(cudaMemcpyAsync(o.b, b_h, sizeof(double) * (size_t)nrows, cudaMemcpyHostToDevice, cs));
    MPG_CUDA_H(cudaMemcpyAsync(o.x, x_h, sizeof(double) * (size_t)nrows, cudaMemcpyHostToDevice, cs));
    MPG_CUDA_H(cudaMemcpyAsync(o.inds, inds_h, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, cs));
    MPG_CUDA_H(cudaEventRecord(e_inds, cs));

    // cast threads: chunk c of the fp32 operator
    const int want = ctx->tune.host_threads > 0 ? ctx->tune.host_threads : std::min(host_cpus(), 32);
    const int nthreads = (int)std::max<int64_t>(1, std::min<int64_t>(want, nch));
    for (int t = 0; t < nthreads; ++t)
        workers.emplace_back([&]() {
            for (;;) {
                const int64_t c = next.fetch_add(1, std::memory_order_relaxed);
                if (c >= nch || abandon.load(std::memory_order_relaxed)) return;
                const int64_t lo = c * CH, len = std::min(CH, nnz - lo);
                cast_range(vals64_h + lo, stage + lo, (size_t)len);
                done[(size_t)c].store(1, std::memory_order_release);
            }
        });

    // feeder: finished fp32 chunks first, fp64 chunks in the gaps; at most two copies queued so that a chunk that becomes ready is on
    // the wire within one chunk time
    const int device = ctx->device;
    feeder = std::thread([&, device]() {
        bool ok = cudaSetDevice(device) == cudaSuccess;
        int64_t i32 = 0, i64 = 0, queued = 0;
        while (ok && i32 < nch) {
            if (abandon.load(std::memory_order_relaxed)) { ok = false; break; }
            if (done[(size_t)i32].load(std::memory_order_acquire)) {
                const int64_t lo = i32 * CH, len = std::min(CH, nnz - lo);
                ok = cudaMemcpyAsync(o.vals32 + lo, stage + lo, sizeof(float) * (size_t)len, cudaMemcpyHostToDevice, cs) == cudaSuccess;
                ++i32;
            } else if (i64 < nch) {
                const int64_t lo = i64 * CH, len = std::min(CH, nnz - lo);
                ok = cudaMemcpyAsync(o.vals + lo, vals64_h + lo, sizeof(double) * (size_t)len, cudaMemcpyHostToDevice, cs) == cudaSuccess;
                ++i64;
            } else {
                std::this_thread::yield();
                continue;
            }
            if (ok && queued >= 2) ok = cudaEventSynchronize(ring[queued & 1]) == cudaSuccess;   // the copy queued two back has finished
            if (ok) ok = cudaEventRecord(ring[queued & 1], cs) == cudaSuccess;
            ++queued;
        }
        if (ok) ok = cudaEventRecord(e_v32, cs) == cudaSuccess;
        v32_recorded.store(ok ? 1 : -1, std::memory_order_release);
        if (ok && i64 < nch) {
            const int64_t lo = i64 * CH;
            ok = cudaMemcpyAsync(o.vals + lo, vals64_h + lo, sizeof(double) * (size_t)(nnz - lo), cudaMemcpyHostToDevice, cs) == cudaSuccess;
        }
        if (ok) ok = cudaEventRecord(df.ev, cs) == cudaSuccess;
        if (!ok) (void)cudaGetLastError();
        df.recorded.store(ok ? 1 : -1, std::memory_order_release);
    });

    // meanwhile on this thread: is x0 zero?  then the plan, which only needs row map and indices
    df.x0_zero = all_plus_zero(x_h, (size_t)nrows);
    MPG_CUDA_H(cudaStreamWaitEvent(ctx->stream, e_inds, 0));
    MPG_TRY_H(mpg_csr_create(ctx, nrows, nrows, nnz, o.row_map, o.inds, &A));
    if (ctx->tune.spmv_packed) {   // the packed structure only needs the indices too: built while the fp32 values travel
        const mpg_sell_plan* plan = nullptr;
        MPG_TRY_H(sell_plan_get(ctx, A, &plan));
    }
    int s32;
    while ((s32 = v32_recorded.load(std::memory_order_acquire)) == 0) std::this_thread::yield();
    if (s32 < 0) { cleanup(); return fail(ctx, MPG_ERR_CUDA, "gmres_solve_host: the host-to-device copy of the fp32 values failed"); }
    MPG_CUDA_H(cudaStreamWaitEvent(ctx->stream, e_v32, 0));
    ctx->defer = &df;
    rc = mpg_gmres_solve(ctx, p, A, o.vals, o.vals32, o.b, o.x, st, hist_inner, cap_inner, hist_outer, cap_outer);
    ctx->defer = nullptr;
    if (rc != MPG_OK) { cleanup(); return rc; }
    join_all();
    MPG_CUDA_H(cudaStreamSynchronize(cs));   // the caller may release its buffers when we return, whether the solve needed the fp64 values or not
    MPG_CUDA_H(cudaEventRecord(e2, ctx->stream));
    MPG_CUDA_H(cudaMemcpyAsync(x_h, o.x, sizeof(double) * (size_t)nrows, cudaMemcpyDeviceToHost, ctx->stream));
    MPG_CUDA_H(cudaEventRecord(e3, ctx->stream));
    MPG_CUDA_H(cudaEventSynchronize(e3));
    float t = 0.f;
    cudaEventElapsedTime(&t, e0, e_v32); st->h2d_ms = t;        // until the solve could start
    cudaEventElapsedTime(&t, e0, df.ev); st->h2d_all_ms = t;    // until the last fp64 value had landed (inside the first cycle)
    cudaEventElapsedTime(&t, e2, e3); st->d2h_ms = t;
    st->h2d_bytes = (int64_t)(sizeof(int) * (size_t)(nrows + 1) + sizeof(int) * (size_t)nnz + (sizeof(double) + sizeof(float)) * (size_t)nnz +
                              2 * sizeof(double) * (size_t)nrows);
    st->host_overlap = nthreads;
    cleanup();
    return MPG_OK;
}

}  // namespace

extern "C" int mpg_gmres_solve_host(mpg_ctx* ctx, const mpg_gmres_params* p, int nrows, int64_t nnz, const int* row_map_h, const int* inds_h,
                                    const double* vals64_h, const double* b_h, double* x_h, mpg_gmres_stats* st, double* hist_inner,
                                    int64_t cap_inner, double* hist_outer, int64_t cap_outer) {
    MPG_REQUIRE(ctx, p && row_map_h && inds_h && vals64_h && b_h && x_h && st && nrows >= 0 && nnz >= 0, "gmres_solve_host: bad argument");
    MPG_REQUIRE(ctx, ctx->dist == nullptr, "gmres_solve_host: takes a global matrix; with a communicator attached use mpg_gmres_solve on the local slab");
    const bool overlap = ctx->tune.host_overlap && p->mode == MPG_MODE_MIXED && p->prec != MPG_PREC_ILU_JACOBI && nnz >= ctx->tune.host_overlap_min_nnz && nnz > 0;
    if (overlap) {
        const int rc = solve_host_overlapped(ctx, p, nrows, nnz, row_map_h, inds_h, vals64_h, b_h, x_h, st, hist_inner, cap_inner, hist_outer, cap_outer);
        if (rc != kNeedSerial) return rc;
    }
    return solve_host_serial(ctx, p, nrows, nnz, row_map_h, inds_h, vals64_h, b_h, x_h, st, hist_inner, cap_inner, hist_outer, cap_outer);
}
