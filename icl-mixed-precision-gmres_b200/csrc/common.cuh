// common.cuh — context, error handling and device helpers shared by every translation unit of
// libmpgmres_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "../../include/mpgmres_b200.h"

#define MPG_PROF_NCLASS 8
namespace mpg {

constexpr int kMaxCols = 256;          // widest basis block any fused kernel accepts (restart length + 1)
constexpr int kMaxPartBlocks = 2048;   // upper bound on the grid of a reducing kernel

enum SpmvPart { SPMV_ALL = 0, SPMV_INTERIOR = 1, SPMV_BOUNDARY = 2, SPMV_ORDERED = 3 };   // ORDERED: every slice in ONE launch, interior slices first, the
                                                                                     // CTAs of the boundary slices wait for the halo flags (packed operator)

struct Tuning {
    int spmv_ctas_per_sm = 8;
    int vpass_stages = 0;     // 0 = auto
    int vpass_serpentine = 1; // alternate traversal direction between consecutive V passes (L2 reuse)
    int gemvt_rb = 1;         // register row-block kernel for gemv-T
    int gemvt_rows_per_block = 8192;
    int passA_rb = 0;         // use gemvt_rb for the first CGS pass (h = V'w) instead of the staged vpass kernel
    int vdirect_max_cols_a = 8;   // widest basis for the register-resident fused kernel, pass A (h = V'w)
    int vdirect_max_cols_b = 8;   //   ... and pass B (w -= V h; c = V'w)
    int vrow_max_cols_a = 32;     // pass A: widest basis for the row-owner staged kernel
    int vrow_max_cols = 56;   // pass B: widest basis handled by the row-owner staged kernel (vrow); wider ones take the column-owner vpass
    int fuse_min_cols = 1;    // basis width from which the fused 3-pass kernels are used (below: gemv-T / gemv-N register kernels, 4 passes)
    int cgs2_fused = 1;       // fused update + gemv-T second pass (3 passes) vs separate gemv-N, gemv-T (4 passes)
    int gemvn_ctas_per_sm = 4;
    int red_ctas_per_sm = 4;
    int dist_peer_halo = 1;   // multi-GPU halo exchange by stores into the neighbours' memory (0: pack + ncclSend/ncclRecv)
    int fuse_tail = 1;        // solver: normalisation of the new basis vector and the Givens update of the column in one launch
    int use_pdl = 1;          // programmatic dependent launch for the kernels of the Arnoldi loop: 0 off, 1 on one GPU when the operand has at
                              // most pdl_max_rows rows, 2 always; off while profiling.  Measured gain (profiles/r02i_pdl_threshold.txt): lap2d 512^2
                              // fp64 m=50 +12 %, cd27 64^3 m=100 0 %, lap2d 1024^2 m=50 +4 %, cd27 100^3 m=100 -5 %, cd27 128^3 -3 %, 16.7 M rows
                              // -8 %; with a communicator attached -6 % .. -21 % at 1-2 M rows per rank
    int pdl_max_rows = 3000000;   // with the late trigger (pdl_trigger_early): +8 % at 0.26 M rows, +4 % at 1 M, +1 % at 2.1 M, 0 at 4.2 M, -2 % at 16.7 M
    int spmv_packed = 1;      // solver: run the inner SpMV on the packed (sliced-ELL) copy of the matrix when it packs well (sell.cu)
    int dist_overlap = 1;     // multi-GPU SpMV: rows without halo columns run between the halo push and the wait for the neighbours' data
    int dist_peer_reduce = 1; // multi-GPU reductions inside the kernels over peer memory (0: NCCL all-reduce + epilogue kernel)
    int residual_packed = 1;  // solver: fp64 outer residual r = b - A x on a packed fp64 copy of the matrix (0: CSR kernel on the caller's arrays)
    int values_static = 0;    // solver: 1 = the caller promises not to change matrix VALUES between solves on the same mpg_csr / value pointers,
                              // so the packed copies are built once (the reference builds SparseMatrix<float>(A) once, outside its solve timer)
    int sell_lpt = 1;         // SELL-C-sigma plans, launch order of the slices (read when the plan is built): 0 window order, 1 slices that would miss
                              // their start deadline in window order move to the front, 2 all slices longest first (sell.cu sell_plan_get)
    int spmv_sigma = 1;       // packed operator: sort rows by length inside windows (SELL-C-sigma) when the plain slices pad too much
    int mgs_fused = 1;        // MGS: pairwise fused passes (w -= h_j v_j ; h_{j+1} = v_{j+1}.w in one kernel) instead of k+1 x {dot, naxpy}
    int dist_fuse_halo = 1;   // multi-GPU: halo gather-and-push rides in the Arnoldi tail kernel, the wait in the boundary-slice SpMV
    int dist_ll_reduce = 1;   // in-kernel all-reduce: flag-in-data mailbox words (one one-way NVLink store per value) instead of data + fence + flag
    int dist_push_in_spmv = 1; // fused halo, stencil-like halos: the push CTAs sit at the head of the SpMV kernel that consumes the column (0: in the
                              // Arnoldi tail that produces it)
    int dist_spmv_one_launch = 1;   // fused halo: interior and boundary slices in ONE launch (boundary CTAs last, they wait for the flags);
                              // 0: two launches (measured on 2.1 M-row slabs: the second launch + its ramp cost ~10 us per iteration)
    int spin_limit_ms = 20000; // multi-GPU: a device-side wait on a peer gives up after this long and raises the context's error word
    int sell_variant = -1;    // packed SpMV kernel variant: bit 0 = x gathers bypass L1, bit 1 = 4 groups per step; -1 = chosen from the plan
    int sell_block = 0;       // packed SpMV threads per CTA (0 = 256)
    int trace = 0;            // 1: host wall-clock of the set-up phases of every solve on stderr (diagnostics)
    int lookahead = 0;        // residual-driven restart policies: speculative Arnoldi steps in flight (0 = auto from a bandwidth estimate)
    int host_overlap = 1;     // mpg_gmres_solve_host, mixed precision: host threads cast the values to fp32 while the indices travel, the fp32
                              // operator goes first and the fp64 values land during the first restart cycle (0: one serial H2D of everything)
    int host_threads = 0;     // host_overlap: cast threads (0 = the CPUs this process may run on, at most 32)
    int host_overlap_min_nnz = 4000000;   // below this the serial copy is a few hundred microseconds: not worth the threads
};

// mpg_gmres_solve_host with host_overlap: the fp64 value array of a mixed-precision solve is still on its way (copy stream) when the
// solve starts.  The feeder thread records `ev` on the copy stream after the last fp64 chunk and then sets `recorded` (1, or -1 after a
// failed copy); the solver waits for it the first time it needs the fp64 operator.
struct DeferredV64 {
    std::atomic<int> recorded{0};
    cudaEvent_t ev = nullptr;
    bool x0_zero = false;     // the caller verified x0 == 0 on the host: r0 = b - A*0 = b needs no operator at all
};

}  // namespace mpg

struct mpg_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string last_error;
    int64_t launches = 0;
    mpg::Tuning tune;

    // reduction scratch: partials[kMaxPartBlocks][kMaxCols + 8] doubles + a ticket counter
    double* partials = nullptr;
    unsigned int* ticket = nullptr;
    // small device scalars (results of host-return dot/nrm2, residual, 1/norm ...)
    double* dscal = nullptr;   // 1024 doubles: [0,64) scalars, [64,64+264) coefficient scratch of mpg_add_vector_*
    double* hscal = nullptr;   // pinned host mirror, 64 doubles
    int vpass_parity = 0;      // serpentine direction toggle

    // per-kernel-class device timers (CUDA events on the launching stream), off by default
    bool prof_on = false;
    struct ProfRec { int cls; cudaEvent_t a, b; double bytes; };
    std::vector<ProfRec> prof_pending;
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[MPG_PROF_NCLASS] = {0};
    double prof_bytes[MPG_PROF_NCLASS] = {0};
    int64_t prof_launches[MPG_PROF_NCLASS] = {0};

    // grow-only device staging arena of the host-buffer entry point (no cudaMalloc/cudaFree per call)
    void* arena = nullptr;
    size_t arena_bytes = 0;
    // host_overlap (hostpath.cu): copy stream, grow-only pinned staging for the fp32 values the host threads produce, and the
    // hand-over of the late fp64 values to the running solve
    cudaStream_t copy_stream = nullptr;
    float* stage32 = nullptr;
    size_t stage32_elems = 0;
    mpg::DeferredV64* defer = nullptr;

    // multi-GPU (dist.cu): communicator + partition attached to this context, raw reduction buffer
    struct mpg_dist* dist = nullptr;
    double* red_raw = nullptr;   // kMaxCols + 8 doubles

    // error word written by device-side waits that gave up (bounded spins on a peer GPU): mapped pinned host memory, so the
    // host can read it without synchronising; 0 = fine
    unsigned int* dev_err = nullptr;        // host pointer
    unsigned int* dev_err_d = nullptr;      // device alias

    // optional per-CTA phase timestamps of the staged V-pass kernels (mpg_debug_timing, tools/vpass_timeline.py)
    unsigned long long* dbg = nullptr;

    // TMA tensor-map cache of the staged V-pass kernels (ortho.cu): per context, never shared between contexts
    void* ortho_cache = nullptr;
    void (*ortho_cache_free)(void*) = nullptr;

    // cached solver workspace (see solver.cu)
    void* ws = nullptr;
    void (*ws_free)(void*) = nullptr;
};

struct mpg_sell_plan;   // sell.cu: packed (sliced-ELL) structure of a matrix
struct mpg_ilu_plan;    // ilu.cu: diagonal positions + level schedule of the ILU(0) factorisation

struct mpg_csr {
    int nrows = 0, ncols = 0;
    int64_t nnz = 0;
    const int* row_map = nullptr;  // not owned
    const int* inds = nullptr;     // not owned
    // SpMV plan (owned): nnz-split tiles
    int tile_nnz = 0;
    int ntiles = 0;
    int* tile_row = nullptr;       // [ntiles+1] row containing the first nonzero of each tile
    void* carry = nullptr;         // [2*ntiles] doubles: carry_in / carry_out partial row sums
    int device = 0;
    int has_empty_rows = 0;        // non-canonical input: SpMV takes the warp-per-row kernel
    // local slab of a partitioned matrix (ncols > nrows, columns >= nrows are halo slots): tiles that touch no halo
    // column first, then the others - lets the solver run the former while the halo is in flight
    int* tile_list = nullptr;      // [ntiles] or null
    int n_interior_tiles = 0;
    // packed structure (shared by the fp32 and fp64 value arrays), built on first use (sell.cu)
    mpg_sell_plan* sell = nullptr;
    int sell_tried = 0;
    mpg_ilu_plan* ilu = nullptr;   // built on first factorisation (ilu.cu)
};

namespace mpg {

void sell_plan_free(mpg_sell_plan* p);
void ilu_plan_free(mpg_ilu_plan* p);

// Plan / packed-matrix / per-solve temporaries come from the device's stream-ordered memory pool (release threshold raised in
// mpg_ctx_create) through a small per-device block cache (blas1.cu): pool_free() parks blocks of >= 1 MiB and pool_alloc() hands a
// parked block of fitting size back, so that a caller who rebuilds plans and packed copies for every solve (the host-buffer entry
// point, a new mpg_csr per call) reaches a steady state with NO driver allocation at all.  (Measured with peer access enabled on a
// 2-GPU run: free + re-allocation of ~4 GB of plan arrays through the driver stalled 0.2-0.9 s every few solves.)
cudaError_t pool_alloc(mpg_ctx* ctx, void** p, size_t bytes);
void pool_free(void* p);          // waits for the stream the block was handed out on (see blas1.cu); accepts null and blocks that did not come from pool_alloc
void pool_stream_changed();       // a context switched streams: blocks handed out before are freed with a device-wide wait
void pool_trim(int device);       // releases every parked block of the device
template <class P> inline cudaError_t pool_alloc(mpg_ctx* ctx, P** p, size_t bytes) { return pool_alloc(ctx, reinterpret_cast<void**>(p), bytes); }

inline int fail(mpg_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->last_error = msg;
    return code;
}

#define MPG_CUDA(ctx, expr)                                                                             \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) (void)cudaGetLastError(); /* do not leave a stale error for later calls */ \
        if (_e != cudaSuccess)                                                                          \
            return mpg::fail(ctx, MPG_ERR_CUDA,                                                         \
                             std::string(#expr) + ": " + cudaGetErrorString(_e) + " @" + __FILE__ + ":" + \
                                 std::to_string(__LINE__));                                             \
    } while (0)

#define MPG_CHECK_LAUNCH(ctx)                                                                       \
    do {                                                                                            \
        (ctx)->launches++;                                                                          \
        cudaError_t _e = cudaGetLastError();                                                        \
        if (_e != cudaSuccess)                                                                      \
            return mpg::fail(ctx, MPG_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(_e) + \
                                                    " @" + __FILE__ + ":" + std::to_string(__LINE__)); \
    } while (0)

// kernel<<<grid, block, smem, ctx->stream>>>(args...) with the programmatic-stream-serialization attribute
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(mpg_ctx* ctx, int64_t rows, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    // auto (1): single GPU and small operands only.  Measured in round 2 (profiles/r02h_*): at 2.1 M rows PDL costs 3 % on one GPU and
    // 6 % (stencil) to 21 % (all-to-all halo) with a communicator attached - early-resident dependents sit on the SMs while the last
    // CTA of a reducing kernel talks to the peers
    const bool on = ctx->tune.use_pdl == 2 || (ctx->tune.use_pdl == 1 && rows <= ctx->tune.pdl_max_rows && ctx->dist == nullptr);
    cfg.numAttrs = (on && !ctx->prof_on) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

#define MPG_REQUIRE(ctx, cond, msg)                                                                                  \
    do {                                                                                                             \
        if ((ctx) == nullptr) return MPG_ERR_ARG; /* a null (closed) context is an argument error, never a crash */    \
        if (!(cond)) return mpg::fail(ctx, MPG_ERR_ARG, (msg));                                                      \
    } while (0)

#define MPG_TRY(expr)              \
    do {                           \
        int _rc = (expr);          \
        if (_rc != MPG_OK) return _rc; \
    } while (0)

inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
// diagnostics (knob `trace`): synchronise and print the host wall-clock since the previous mark
struct Trace {
    mpg_ctx* ctx;
    double t0;
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
    explicit Trace(mpg_ctx* c) : ctx(c), t0(c->tune.trace ? now() : 0.0) {}
    void mark(const char* what) {
        if (!ctx->tune.trace) return;
        cudaStreamSynchronize(ctx->stream);
        const double t = now();
        fprintf(stderr, "[mpg trace] %-28s %9.3f ms\n", what, t - t0);
        t0 = t;
    }
};
int check_dev_err(mpg_ctx* ctx);   // blas1.cu

// dist.cu: all-reduce `count` raw sums over the ranks and apply the epilogue (no-op when no communicator is attached)
int dist_finish_reduction(mpg_ctx* ctx, const struct Epi& e, int count, int tbytes);
// epilogue descriptor for the next reducing kernel of this context: plain on one GPU; with a communicator attached it
// carries either the peer-memory mailboxes (default) or the raw buffer for the NCCL path
struct Epi make_epi(mpg_ctx* ctx, int kind, void* p0, void* p1, double alpha, double beta);

// RAII timer for one kernel (or a kernel + its fix-up) of class `cls`, carrying its ALGORITHMIC bytes
// (compulsory traffic).  No-op unless mpg_prof_enable(ctx, 1).
struct ProfScope {
    mpg_ctx* ctx;
    cudaEvent_t a = nullptr, b = nullptr;
    int cls;
    double bytes;
    static cudaEvent_t get(mpg_ctx* c) {
        if (!c->prof_pool.empty()) { cudaEvent_t e = c->prof_pool.back(); c->prof_pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
    ProfScope(mpg_ctx* c, int cls_, double bytes_) : ctx(c), cls(cls_), bytes(bytes_) {
        if (!ctx->prof_on) return;
        a = get(ctx); b = get(ctx);
        cudaEventRecord(a, ctx->stream);
    }
    ~ProfScope() {
        if (!a) return;
        cudaEventRecord(b, ctx->stream);
        ctx->prof_pending.push_back({cls, a, b, bytes});
    }
};

// ---- device helpers ------------------------------------------------------------------------------------
template <class T> struct Vec4;
template <> struct Vec4<float> { using type = float4; };
template <> struct Vec4<double> { using type = double4; };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Programmatic dependent launch (sm_90+): kernels of the Arnoldi loop are launched with the programmatic-stream-
// serialization attribute (launch_pdl below).  pdl_trigger() lets the NEXT kernel's CTAs be scheduled as soon as SM
// resources free up, so its launch latency and prologue overlap this kernel's tail; pdl_wait() blocks until the
// PREVIOUS kernel has completed and its memory is visible - it must precede the first global-memory access.  Both are
// no-ops for kernels launched the ordinary way.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// Where a kernel lets its dependents be scheduled.  Small operands (launch-bound kernels of 5-15 us): at entry, so the whole launch
// latency of the next kernel hides behind this one (+12 % on the 262 K-row fp64 config).  Large operands: only when the CTA has done
// its streaming work (pdl_trigger() at that point) - dependents that become resident while this kernel still streams take shared
// memory and registers away from it (a 200 KB V-pass CTA next to SpMV CTAs shrinks their L1) and cost 3-8 %; released late they
// still hide their launch latency and prologue behind the last-CTA reduction and the cross-GPU combine of this kernel.
// Scalars and coefficient vectors that the PREVIOUS kernel of the stream produced (h, c, 1/norm, Givens state): loaded around L1
// (ld.global.cg).  A dependent launched programmatically shares SMs - and their L1 - with its still-running primary; a line the
// primary's CTAs read early (h(j,k)) can hold the stale neighbour (h(j+1,k)) that the primary's last CTA writes at the end, and the
// read-only path (ld.global.nc) is only defined for data nobody writes while the kernel runs.
template <class T> __device__ __forceinline__ T ld_fresh(const T* p) { return __ldcg(p); }
constexpr int64_t kPdlEarlyRows = 600000;
__device__ __forceinline__ void pdl_trigger_early(int64_t rows) { if (rows <= kPdlEarlyRows) pdl_trigger(); }

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// streaming (read-once) loads: bypass L1 allocation so L1 stays available for gathered vectors
__device__ __forceinline__ int4 ldg_stream(const int4* p) {
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ double2 ldg_stream(const double2* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ double ldg_stream(const double* p) {
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ int ldg_stream(const int* p) {
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// Deterministic grid-wide reduction of `ncols` running sums.
//   1. every block calls block_reduce_store() to put its per-column partial (double) in
//      partials[blockIdx.x * ld + j];
//   2. grid_last_block() returns true in exactly one block — the last to arrive — after all partials are
//      visible; that block sums them in a FIXED order (column per warp, blocks strided over lanes, xor
//      tree) so the result does not depend on block scheduling.
__device__ __forceinline__ bool grid_last_block(unsigned int* ticket) {
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
        if (is_last) *ticket = 0u;  // re-arm for the next launch (stream-ordered)
    }
    __syncthreads();
    if (is_last) __threadfence();
    return is_last;
}

// Epilogue of a grid-wide reduction: what happens to the reduced value(s).  On one GPU the last CTA applies it in the
// reducing kernel itself; with a communicator attached (dist.cu) the kernel stores the raw local sums instead, they are
// all-reduced over the ranks (fp64), and epilogue_kernel applies the same function afterwards.
// In-kernel all-reduce over NVLink peer memory (multi-GPU, dist.cu): every rank owns a mailbox
// [slot][source rank][kMaxCols+8] doubles + one 64-bit flag per (slot, source); the peers' mailboxes are mapped into
// this process (CUDA IPC).  The last CTA of a reducing kernel pushes its local sums into every rank's mailbox with
// plain stores over NVLink, publishes a sequence number (release, system scope), waits for the other ranks' flags
// (acquire) and adds the contributions in rank order - identical bits on every rank - before running the epilogue.
// One kernel = local reduction + cross-GPU combine + epilogue; no NCCL launch, no separate epilogue launch.
constexpr int kMaxPeers = 8;
constexpr int kMboxSlots = 4;
constexpr int kMboxStride = 256 + 8;
struct PeerComm {
    int world = 0;   // 0: disabled
    int rank = 0;
    unsigned long long seq = 0;
    double* mbox[kMaxPeers];
    unsigned long long* flag[kMaxPeers];
    ulonglong2* ll[kMaxPeers];              // flag-in-data mailbox: every 8-byte word carries 32 bits of payload and the 32-bit sequence number
    int use_ll = 0;
    unsigned int* err = nullptr;            // device error word (mpg_ctx::dev_err_d)
    unsigned long long spin_limit_ns = 0;   // 0: wait for ever
};

// Bounded wait on a flag another GPU publishes: spins on an acquire load until *flag >= seq; after spin_limit_ns it gives up,
// raises the context's error word (the host reports MPG_ERR_STATE at its next synchronisation) and lets the kernel finish
// with whatever data is there - a dead or diverged peer no longer hangs every GPU of the job.
enum { DEV_ERR_REDUCE_TIMEOUT = 1, DEV_ERR_HALO_TIMEOUT = 2 };
__device__ __forceinline__ void wait_flag(const unsigned long long* flag, unsigned long long seq, unsigned long long limit_ns, unsigned int* err, unsigned int code) {
    unsigned long long v;
    unsigned long long t0 = 0;
    unsigned int spins = 0;
    for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
        if (v >= seq) return;
        if (limit_ns && (++spins & 1023u) == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > limit_ns) {
                if (err) atomicOr(err, code);
                return;
            }
        }
    }
}

// ---- halo exchange over NVLink peer memory (push model, dist.cu) --------------------------------------------------------
// The sender gathers the rows a neighbour needs and stores them straight into that neighbour's memory - its inbox, or (fused
// path) the halo tail of the neighbour's own copy of the Krylov basis column - then publishes the exchange number in the
// neighbour's flag word; kPushBlocksPerPeer CTAs per neighbour (one SM cannot keep an NVLink busy with stores).
constexpr int kPushBlocksMin = 32, kPushBlocksMax = 512;   // CTAs per neighbour: one SM cannot keep an NVLink busy with stores
struct PushArgs {
    int npeers = 0;
    int bpp = kPushBlocksMin;               // CTAs per neighbour of this launch (push_blocks())
    const int* send_idx[kMaxPeers];
    long long count[kMaxPeers];
    void* dst[kMaxPeers];                   // where our rows go in the neighbour's memory
    unsigned long long* flag[kMaxPeers];    // the neighbour's flag for (slot, this rank)
    unsigned long long seq = 0;
    unsigned int* counters = nullptr;       // kMaxPeers arrival counters (device)
};
// 32 CTAs move a stencil halo (one 256^2 plane = 65 536 entries) in a few microseconds; an all-to-all halo of millions of
// entries (power-law columns) needs the whole machine: ~4096 entries per CTA, capped
inline int push_blocks(long long max_count) {
    const long long b = (max_count + 4095) / 4096;
    return (int)(b < kPushBlocksMin ? kPushBlocksMin : (b > kPushBlocksMax ? kPushBlocksMax : b));
}
// block `pb` (0 <= pb < npeers * bpp) of a push: dst[i] = scale * x[idx[i]] (scale == null: plain copy); 4 gathers in flight per thread
template <class T>
__device__ __forceinline__ void halo_push_block(const PushArgs& a, int pb, const T* __restrict__ x, const T* __restrict__ scale) {
    const int q = pb / a.bpp, part = pb % a.bpp;
    T* dst = static_cast<T*>(a.dst[q]);
    const int* idx = a.send_idx[q];
    const long long cnt = a.count[q];
    const T al = scale ? ld_fresh(scale) : T(1);
    const long long stride = (long long)a.bpp * blockDim.x;
    long long i = (long long)part * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < cnt; i += 4 * stride) {
        const int j0 = ldg_stream(idx + i), j1 = ldg_stream(idx + i + stride), j2 = ldg_stream(idx + i + 2 * stride), j3 = ldg_stream(idx + i + 3 * stride);
        const T v0 = x[j0], v1 = x[j1], v2 = x[j2], v3 = x[j3];
        if (scale) { dst[i] = al * v0; dst[i + stride] = al * v1; dst[i + 2 * stride] = al * v2; dst[i + 3 * stride] = al * v3; }
        else { dst[i] = v0; dst[i + stride] = v1; dst[i + 2 * stride] = v2; dst[i + 3 * stride] = v3; }
    }
    for (; i < cnt; i += stride) dst[i] = scale ? al * x[idx[i]] : x[idx[i]];
    // ONE system-scope fence per CTA, by thread 0 after the barrier (fences are cumulative: the barrier makes the other threads'
    // remote stores happen-before it).  A fence in every warp costs an NVLink round trip each and they queue up per SM: with
    // thousands of push CTAs (all-to-all halos) that was most of the kernel.
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        // the last block of this neighbour publishes the exchange number (the counter chain + its fence order all stores before it)
        if (atomicAdd(a.counters + q, 1u) == (unsigned)a.bpp - 1u) {
            a.counters[q] = 0u;
            __threadfence_system();
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.flag[q]), "l"(a.seq) : "memory");
        }
    }
}

// halo of an SpMV input that neighbour GPUs push into this GPU's memory (dist.cu): the flags to wait for
struct HaloWait {
    int npeers = 0;
    int wait_from = 0;                      // only CTAs that hold a slice position >= wait_from wait (SPMV_ORDERED: the interior slices come first)
    const unsigned long long* flag[kMaxPeers];
    unsigned long long seq = 0;
    unsigned int* err = nullptr;
    unsigned long long spin_limit_ns = 0;
};
__device__ __forceinline__ void halo_wait_block(const HaloWait& hw) {
    if ((int)threadIdx.x < hw.npeers) wait_flag(hw.flag[threadIdx.x], hw.seq, hw.spin_limit_ns, hw.err, DEV_ERR_HALO_TIMEOUT);
    __syncthreads();
}

enum EpiKind { EPI_DOT = 0, EPI_NRM2 = 1, EPI_COEF = 2, EPI_COEF_ACCUM = 3, EPI_NORM_INV = 4, EPI_GEMVT = 5, EPI_MAX = 6 };
struct Epi {
    int kind;
    void* p0;      // DOT/NRM2/MAX: out | COEF*: coef_out | NORM_INV: norm_out | GEMVT: y
    void* p1;      // COEF*: hcol | NORM_INV: inv_out
    double alpha, beta;
    double* raw;   // non-null: store the raw fp64 sums here and skip the epilogue (multi-GPU through NCCL)
    PeerComm peer; // world > 1: combine over peer memory inside the kernel (multi-GPU, default)
};
template <class T>
__device__ __forceinline__ void apply_epi(const Epi& e, int j, double s) {
    if (e.raw) { e.raw[j] = s; return; }
    T* p0 = static_cast<T*>(e.p0);
    T* p1 = static_cast<T*>(e.p1);
    switch (e.kind) {
        case EPI_DOT: p0[j] = (T)s; break;
        case EPI_MAX: p0[j] = (T)s; break;
        case EPI_NRM2: p0[j] = (T)sqrt(s); break;
        case EPI_COEF: { const T c = (T)s; p0[j] = c; if (p1 && p1 != p0) p1[j] = c; } break;
        case EPI_COEF_ACCUM: { const T c = (T)s; p0[j] = c; p1[j] = p1[j] + c; } break;   // axpy(1, weights, h_col) Orthogonalization.hpp:133
        case EPI_NORM_INV: { const T nrm = (T)sqrt(s); p0[j] = nrm; p1[j] = T(1) / nrm; } break;   // Orthogonalization.hpp:55,59
        case EPI_GEMVT: { const T a = (T)e.alpha, b = (T)e.beta; p0[j] = (b == T(0)) ? a * (T)s : fma(a, (T)s, b * p0[j]); } break;
    }
}

// sum over blocks of partials[b*ld + j] for the calling warp's column j (all lanes get the result)
__device__ __forceinline__ double reduce_partials_column(const double* partials, int ld, int nblocks, int j) {
    const int lane = threadIdx.x & 31;
    double acc = 0.0;
    for (int b = lane; b < nblocks; b += 32) acc += __ldcg(partials + (size_t)b * ld + j);
    return warp_sum(acc);
}

// Called by ALL threads of the last CTA with the `count` locally reduced values in shared memory: cross-GPU combine
// over peer memory when a PeerComm is attached, then the epilogue.
template <class T>
__device__ __forceinline__ void finish_reduction(const Epi& e, int count, double* red_s) {
    if (e.peer.world > 1 && e.peer.use_ll) {
        // Flag-in-data exchange (the idea of NCCL's LL protocol): a double travels as two naturally aligned 8-byte words, each = 32 bits
        // of payload + the 32-bit sequence number of this reduction.  8-byte stores are single-copy atomic, so a word whose tag matches
        // is complete: no fence, no separate flag, no wait for NVLink write acknowledgements on the sender - the latency of a reduction
        // is ONE one-way store instead of round trip + flag + read.  Ranks are at most one reduction apart (nobody finishes reduction s
        // before everybody contributed to it), so kMboxSlots >= 2 slots never see a live overwrite; stale words carry older tags.
        const int P = e.peer.world, r = e.peer.rank;
        const int slot = (int)(e.peer.seq % kMboxSlots);
        const unsigned long long tag = (e.peer.seq & 0xffffffffull) << 32;
        for (int idx = threadIdx.x; idx < count * P; idx += blockDim.x) {
            const int q = idx / count, j = idx - q * count;
            const unsigned long long bits = (unsigned long long)__double_as_longlong(red_s[j]);
            ulonglong2* dst = e.peer.ll[q] + ((size_t)slot * P + r) * kMboxStride + j;
            asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"((bits & 0xffffffffull) | tag), "l"((bits >> 32) | tag) : "memory");
        }
        __syncthreads();   // red_s is overwritten below
        for (int j = threadIdx.x; j < count; j += blockDim.x) {
            const ulonglong2* src = e.peer.ll[r] + (size_t)slot * P * kMboxStride + j;
            double t = 0.0;
            for (int q = 0; q < P; ++q) {
                unsigned long long w0, w1, t0 = 0;
                unsigned int spins = 0;
                for (;;) {
                    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src + (size_t)q * kMboxStride) : "memory");
                    if ((w0 & 0xffffffff00000000ull) == tag && (w1 & 0xffffffff00000000ull) == tag) break;
                    if (e.peer.spin_limit_ns && (++spins & 1023u) == 0) {
                        unsigned long long now;
                        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
                        if (t0 == 0) t0 = now;
                        else if (now - t0 > e.peer.spin_limit_ns) { if (e.peer.err) atomicOr(e.peer.err, DEV_ERR_REDUCE_TIMEOUT); break; }
                    }
                }
                const double c = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
                t = (q == 0) ? c : ((e.kind == EPI_MAX) ? fmax(t, c) : t + c);
            }
            red_s[j] = t;
        }
        __syncthreads();
    } else if (e.peer.world > 1) {
        const int P = e.peer.world, r = e.peer.rank;
        const int slot = (int)(e.peer.seq % kMboxSlots);
        for (int idx = threadIdx.x; idx < count * P; idx += blockDim.x) {
            const int q = idx / count, j = idx - q * count;
            volatile double* dst = e.peer.mbox[q] + ((size_t)slot * P + r) * kMboxStride + j;
            *dst = red_s[j];
        }
        // no fence in every warp here: the barrier orders the CTA's mailbox stores before the release stores below, and a release is
        // cumulative - one system-scope fence (inside st.release.sys, warp 0 only) instead of one NVLink round trip per warp on the
        // critical path of every reduction
        __syncthreads();
        if ((int)threadIdx.x < P) {
            unsigned long long* f = e.peer.flag[threadIdx.x] + slot * P + r;
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(e.peer.seq) : "memory");
            const unsigned long long* mine = e.peer.flag[r] + slot * P + threadIdx.x;
            wait_flag(mine, e.peer.seq, e.peer.spin_limit_ns, e.peer.err, DEV_ERR_REDUCE_TIMEOUT);
        }
        __syncthreads();
        for (int j = threadIdx.x; j < count; j += blockDim.x) {
            const double* src = e.peer.mbox[r] + (size_t)slot * P * kMboxStride + j;
            double t = __ldcv(src);
            for (int q = 1; q < P; ++q) {
                const double c = __ldcv(src + (size_t)q * kMboxStride);
                t = (e.kind == EPI_MAX) ? fmax(t, c) : t + c;
            }
            red_s[j] = t;
        }
        __syncthreads();
    }
    for (int j = threadIdx.x; j < count; j += blockDim.x) apply_epi<T>(e, j, red_s[j]);
}

// last CTA: fixed-order sum of the per-CTA partials of `count` columns, then finish_reduction.
// Two shapes, chosen from (count, blockDim) only, so the summation order is a function of the launch configuration:
//   * few columns: one warp per column, the blocks strided over the lanes, xor tree;
//   * many columns (more than two per warp) and a scratch buffer of >= blockDim doubles in shared memory: thread
//     (r, j) sums column j over the blocks r, r+R, ... with coalesced loads kept 8 deep in flight, then the R partial
//     sums are added in r order.  (The warp-per-column shape serialises count/nwarps dependent L2 round trips: 22 us
//     at 100 columns, tools/vpass_timeline.py.)
template <class T>
__device__ __forceinline__ void last_block_finish(const Epi& e, const double* partials, int ldp, int nblocks, int count, double* scratch = nullptr) {
    __shared__ double red_s[kMaxCols + 8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int R = scratch ? min((int)blockDim.x / max(count, 1), 16) : 0;
    if (count > 2 * nw && R >= 2) {
        const int r = threadIdx.x / count, j = threadIdx.x - r * count;
        if (r < R) {
            double acc = 0.0;
#pragma unroll 16
            for (int b = r; b < nblocks; b += R) acc += __ldcg(partials + (size_t)b * ldp + j);   // same order, 16 L2 loads in flight
            scratch[r * count + j] = acc;
        }
        __syncthreads();
        for (int jj = threadIdx.x; jj < count; jj += blockDim.x) {
            double t = scratch[jj];
            for (int q = 1; q < R; ++q) t += scratch[q * count + jj];
            red_s[jj] = t;
        }
    } else {
        for (int j = wid; j < count; j += nw) {
            const double sred = reduce_partials_column(partials, ldp, nblocks, j);
            if (lane == 0) red_s[j] = sred;
        }
    }
    __syncthreads();
    finish_reduction<T>(e, count, red_s);
}

}  // namespace mpg
