// dist.cu — 1-D row-partitioned multi-GPU layer (one process per GPU).  New functionality: the reference is single
// device.  Rank r owns rows [floor(r n/P), floor((r+1) n/P)); every vector on the path is split by
// the same row map; h, s, cos, sin and the Givens / least-squares state are replicated, so every rank takes identical
// control decisions.
//   * SpMV: the local slab's columns are renumbered [local | halo]; x lives in a buffer with the halo appended
//     (the tail of each basis column), filled by mpg_halo_exchange: pack kernel (gather of the rows each peer needs)
//     + grouped ncclSend/ncclRecv over NVLink, received straight into the halo tail.
//   * dot / nrm2 / gemv-T / the fused passes: ONE kernel does the local reduction, the cross-GPU combine and the
//     epilogue.  Its last CTA pushes the <= 257 local fp64 sums into every rank's mailbox with stores over NVLink peer
//     memory (CUDA IPC mapped), publishes a sequence number, waits for the peers' and adds the contributions in rank
//     order (common.cuh finish_reduction) - identical bits on all ranks, no collective launch, no epilogue launch.
//     Fallback / comparison path (tuning knob dist_peer_reduce = 0, or more than 8 ranks): the kernel stores the raw
//     sums, ncclAllReduce(sum, fp64) combines them and epilogue_kernel applies the epilogue.
//     gemv-N / axpy / scal / casts are purely local.
// NCCL is loaded at run time (dlopen "libnccl.so.2": the copy torch already mapped), so the library has no link-time
// NCCL dependency and single-GPU use never touches it.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

using namespace mpg;

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl() {
    static NcclApi api;
    if (api.handle) return api;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) return api;
#define MPG_SYM(f) api.f = reinterpret_cast<decltype(api.f)>(dlsym(api.handle, "nccl" #f))
    MPG_SYM(GetUniqueId); MPG_SYM(CommInitRank); MPG_SYM(CommDestroy); MPG_SYM(AllReduce); MPG_SYM(AllGather); MPG_SYM(Send); MPG_SYM(Recv);
    MPG_SYM(GroupStart); MPG_SYM(GroupEnd); MPG_SYM(GetErrorString);
#undef MPG_SYM
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.AllGather && api.Send && api.Recv && api.GroupStart && api.GroupEnd;
    return api;
}

#define MPG_NCCL(ctx, expr)                                                                                              \
    do {                                                                                                                 \
        ncclResult_t _r = (expr);                                                                                        \
        if (_r != ncclSuccess)                                                                                           \
            return mpg::fail(ctx, MPG_ERR_NCCL, std::string(#expr) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(_r) : "nccl error")); \
    } while (0)

template <class T>
__global__ void epilogue_kernel(Epi e, int count) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    const double s = e.raw[j];
    e.raw = nullptr;
    apply_epi<T>(e, j, s);
}

template <class T>
__global__ void pack_kernel(int64_t count, const int* __restrict__ idx, const T* __restrict__ x, T* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = x[idx[i]];
}

// ---- halo exchange over NVLink peer memory (push model) ---------------------------------------------------------------
// Every rank owns an IPC-shared inbox: [2 slots][n_halo] 8-byte cells + one flag per (slot, source rank).  The sender
// gathers the rows a neighbour needs and stores them straight into that neighbour's inbox, then publishes the exchange
// number; the receiver waits for its neighbours' flags and moves the inbox into the halo tail of the SpMV input.
template <class T>
__global__ void __launch_bounds__(256) halo_push_kernel(const __grid_constant__ PushArgs a, const T* __restrict__ x) {
    halo_push_block<T>(a, blockIdx.x, x, nullptr);
}
// fused path, matrices that do not use the packed SpMV: wait for the neighbours' pushes into the basis column (no data to move)
__global__ void halo_wait_kernel(const __grid_constant__ HaloWait hw) { halo_wait_block(hw); }
struct WaitArgs {
    int npeers;
    const unsigned long long* flag[kMaxPeers];   // own flags for (slot, neighbour)
    unsigned int* err;
    unsigned long long spin_limit_ns;
};
template <class T>
__global__ void __launch_bounds__(256) halo_wait_copy_kernel(WaitArgs a, unsigned long long seq, const T* inbox, T* tail, long long n_halo) {
    if ((int)threadIdx.x < a.npeers) wait_flag(a.flag[threadIdx.x], seq, a.spin_limit_ns, a.err, DEV_ERR_HALO_TIMEOUT);
    __syncthreads();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_halo; i += (long long)gridDim.x * blockDim.x) tail[i] = __ldcv(inbox + i);
}

}  // namespace

struct mpg_dist {
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    int device = 0;
    // partition (rows are local; columns [0, n_local) local, [n_local, n_local + n_halo) halo)
    int64_t n_global = 0, n_local = 0, n_halo = 0;
    struct Peer {
        int rank;
        int64_t send_count, send_offset;   // offset into the packed send buffer
        int64_t recv_count, recv_offset;   // offset into the halo tail
        const int* send_idx;               // device, local row indices (ascending), not owned
    };
    std::vector<Peer> peers;
    int64_t send_total = 0;
    int64_t max_send_count = 0;            // longest send list (sizes the push grid)
    void* send_buf = nullptr;              // send_total doubles
    // peer-memory mailboxes for the in-kernel all-reduce (common.cuh PeerComm)
    void* mbox_own = nullptr;              // [kMboxSlots][world][kMboxStride] doubles, then [kMboxSlots][world] u64 flags, then the same shape in 16-byte
                                           // flag-in-data words (common.cuh finish_reduction)
    void* mbox_map[kMaxPeers] = {nullptr}; // every rank's mailbox mapped into this process (own pointer for self)
    bool peer_ready = false;
    unsigned long long seq = 0;
    // peer-memory halo inboxes: [2][n_halo] 8-byte cells, then [2][kMaxPeers] u64 flags
    void* inbox_own = nullptr;
    void* inbox_map[kMaxPeers] = {nullptr};
    int64_t remote_off[kMaxPeers] = {0};     // per peer (index into `peers`): where our rows land in that peer's halo
    int64_t remote_nhalo[kMaxPeers] = {0};   // per peer: that peer's n_halo (slot stride)
    bool halo_ready = false;
    unsigned long long halo_seq = 0;
    unsigned int* push_counters = nullptr;   // kMaxPeers, device
    // owned by the native set-up (mpg_dist_setup): send lists of all neighbours back to back, ascending halo column ids
    // fused halo: every neighbour's Krylov basis (cached solver workspace) mapped into this process, so that basis column k + 1 is
    // pushed straight into the halo tail of the neighbour's own copy of that column
    struct PeerBasis { unsigned char handle[64]; void* base = nullptr; long long ldv = 0, n_local = 0; int tsize = 0; bool valid = false; };
    PeerBasis vpeer[kMaxPeers];              // indexed by RANK
    bool basis_ready = false;
    char* xchg_stage = nullptr;              // device staging of the per-solve hand-shake
    int* send_idx_own = nullptr;
    int* halo_cols_dev = nullptr;            // [n_halo] global column of every halo slot
    size_t inbox_data_bytes() const { return 2 * (size_t)std::max<int64_t>(n_halo, 1) * 8; }
    size_t inbox_bytes() const { return inbox_data_bytes() + sizeof(unsigned long long) * 2 * kMaxPeers; }
    size_t mbox_data_bytes() const { return sizeof(double) * (size_t)kMboxSlots * world * kMboxStride; }
    size_t mbox_flag_bytes() const { return sizeof(unsigned long long) * (size_t)kMboxSlots * world; }
    size_t mbox_ll_bytes() const { return 16 * (size_t)kMboxSlots * world * kMboxStride; }   // flag-in-data words: 16 B per value
    size_t mbox_bytes() const { return mbox_data_bytes() + mbox_flag_bytes() + mbox_ll_bytes(); }
};

extern "C" int mpg_nccl_unique_id(void* id128) {
    if (!id128) return MPG_ERR_ARG;
    if (!nccl().ok) return MPG_ERR_NCCL;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    return nccl().GetUniqueId(static_cast<ncclUniqueId*>(id128)) == ncclSuccess ? MPG_OK : MPG_ERR_NCCL;
}

extern "C" int mpg_dist_create(mpg_ctx* ctx, const void* id128, int rank, int world, mpg_dist** out) {
    MPG_REQUIRE(ctx, id128 && out && world >= 1 && rank >= 0 && rank < world, "dist_create: bad argument");
    if (!nccl().ok) return fail(ctx, MPG_ERR_NCCL, "libnccl.so.2 could not be loaded");
    mpg_dist* d = new mpg_dist();
    d->rank = rank; d->world = world; d->device = ctx->device;
    MPG_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    MPG_NCCL(ctx, nccl().CommInitRank(&d->comm, world, id, rank));
    *out = d;
    return MPG_OK;
}

extern "C" int mpg_dist_destroy(mpg_dist* d) {
    if (!d) return MPG_OK;
    cudaSetDevice(d->device);
    if (d->comm && nccl().ok) nccl().CommDestroy(d->comm);
    for (int q = 0; q < d->world && q < kMaxPeers; ++q) {
        if (d->mbox_map[q] && q != d->rank) cudaIpcCloseMemHandle(d->mbox_map[q]);
        if (d->inbox_map[q] && q != d->rank) cudaIpcCloseMemHandle(d->inbox_map[q]);
    }
    for (int q = 0; q < d->world && q < kMaxPeers; ++q)
        if (d->vpeer[q].base && q != d->rank) cudaIpcCloseMemHandle(d->vpeer[q].base);
    cudaFree(d->xchg_stage);
    cudaFree(d->inbox_own);
    cudaFree(d->send_idx_own);
    cudaFree(d->halo_cols_dev);
    cudaFree(d->push_counters);
    cudaFree(d->mbox_own);
    cudaFree(d->send_buf);
    delete d;
    return MPG_OK;
}

extern "C" int mpg_dist_set_partition(mpg_ctx* ctx, mpg_dist* d, int64_t n_global, int64_t n_local, int64_t n_halo, int npeers,
                                      const int* peer_ranks, const int64_t* send_counts, const int* const* send_idx_dev,
                                      const int64_t* recv_offsets, const int64_t* recv_counts) {
    MPG_REQUIRE(ctx, d && n_global >= 0 && n_local >= 0 && n_halo >= 0 && npeers >= 0, "dist_set_partition: bad argument");
    // a rank without rows would return early from every reducing entry point and never join the cross-rank reductions: refuse it
    MPG_REQUIRE(ctx, d->world == 1 || n_local >= 1, "dist_set_partition: every rank must own at least one row");
    d->n_global = n_global; d->n_local = n_local; d->n_halo = n_halo;
    d->peers.clear();
    d->max_send_count = 0;
    int64_t off = 0;
    for (int i = 0; i < npeers; ++i) {
        MPG_REQUIRE(ctx, peer_ranks[i] >= 0 && peer_ranks[i] < d->world && peer_ranks[i] != d->rank, "dist_set_partition: bad peer rank");
        MPG_REQUIRE(ctx, recv_offsets[i] >= 0 && recv_offsets[i] + recv_counts[i] <= n_halo, "dist_set_partition: halo range out of bounds");
        d->peers.push_back({peer_ranks[i], send_counts[i], off, recv_counts[i], recv_offsets[i], send_idx_dev[i]});
        d->max_send_count = std::max<int64_t>(d->max_send_count, send_counts[i]);
        off += send_counts[i];
    }
    d->send_total = off;
    cudaFree(d->send_buf);
    d->send_buf = nullptr;
    if (off > 0) MPG_CUDA(ctx, cudaMalloc(&d->send_buf, sizeof(double) * (size_t)off));
    return MPG_OK;
}

// ---- peer-memory mailboxes (CUDA IPC; one process per GPU) ---------------------------------------------------------------
extern "C" int mpg_dist_mailbox_handle(mpg_ctx* ctx, mpg_dist* d, void* handle64) {
    MPG_REQUIRE(ctx, d && handle64, "dist_mailbox_handle: bad argument");
    MPG_REQUIRE(ctx, d->world <= kMaxPeers, "dist_mailbox_handle: at most 8 ranks");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    if (!d->mbox_own) {
        MPG_CUDA(ctx, cudaMalloc(&d->mbox_own, d->mbox_bytes()));
        MPG_CUDA(ctx, cudaMemset(d->mbox_own, 0, d->mbox_bytes()));
    }
    cudaIpcMemHandle_t h;
    MPG_CUDA(ctx, cudaIpcGetMemHandle(&h, d->mbox_own));
    memcpy(handle64, &h, sizeof(h));
    return MPG_OK;
}
// handles: world x 64 bytes, in rank order (all-gathered by the host plumbing)
extern "C" int mpg_dist_open_mailboxes(mpg_ctx* ctx, mpg_dist* d, const void* handles) {
    MPG_REQUIRE(ctx, d && handles && d->mbox_own, "dist_open_mailboxes: bad argument");
    for (int q = 0; q < d->world; ++q) {
        if (q == d->rank) { d->mbox_map[q] = d->mbox_own; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(handles) + (size_t)q * 64, sizeof(h));
        MPG_CUDA(ctx, cudaIpcOpenMemHandle(&d->mbox_map[q], h, cudaIpcMemLazyEnablePeerAccess));
    }
    d->peer_ready = true;
    d->seq = 0;
    return MPG_OK;
}

// halo inbox (after mpg_dist_set_partition: its size is this rank's halo)
extern "C" int mpg_dist_halo_handle(mpg_ctx* ctx, mpg_dist* d, void* handle64) {
    MPG_REQUIRE(ctx, d && handle64 && d->world <= kMaxPeers, "dist_halo_handle: bad argument");
    if (d->inbox_own) { cudaFree(d->inbox_own); d->inbox_own = nullptr; d->halo_ready = false; }
    MPG_CUDA(ctx, cudaMalloc(&d->inbox_own, d->inbox_bytes()));
    MPG_CUDA(ctx, cudaMemset(d->inbox_own, 0, d->inbox_bytes()));
    cudaIpcMemHandle_t h;
    MPG_CUDA(ctx, cudaIpcGetMemHandle(&h, d->inbox_own));
    memcpy(handle64, &h, sizeof(h));
    return MPG_OK;
}
// handles: world x 64 bytes (rank order); per peer of the plan (same order as mpg_dist_set_partition): the offset of our
// rows inside that peer's halo and that peer's halo length
extern "C" int mpg_dist_open_halo(mpg_ctx* ctx, mpg_dist* d, const void* handles, const int64_t* remote_offsets, const int64_t* remote_nhalo) {
    MPG_REQUIRE(ctx, d && handles && d->inbox_own && remote_offsets && remote_nhalo, "dist_open_halo: bad argument");
    MPG_REQUIRE(ctx, (int)d->peers.size() <= kMaxPeers, "dist_open_halo: too many neighbours");
    for (int q = 0; q < d->world; ++q) {
        if (q == d->rank) { d->inbox_map[q] = d->inbox_own; continue; }
        if (d->inbox_map[q]) { cudaIpcCloseMemHandle(d->inbox_map[q]); d->inbox_map[q] = nullptr; }
        bool needed = false;
        for (const auto& p : d->peers) needed = needed || p.rank == q;
        if (!needed) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, static_cast<const char*>(handles) + (size_t)q * 64, sizeof(h));
        MPG_CUDA(ctx, cudaIpcOpenMemHandle(&d->inbox_map[q], h, cudaIpcMemLazyEnablePeerAccess));
    }
    for (size_t i = 0; i < d->peers.size(); ++i) { d->remote_off[i] = remote_offsets[i]; d->remote_nhalo[i] = remote_nhalo[i]; }
    if (!d->push_counters) {
        MPG_CUDA(ctx, cudaMalloc(&d->push_counters, sizeof(unsigned int) * kMaxPeers));
        MPG_CUDA(ctx, cudaMemset(d->push_counters, 0, sizeof(unsigned int) * kMaxPeers));
    }
    d->halo_ready = true;
    d->halo_seq = 0;
    return MPG_OK;
}

// ---- native set-up: partition plan + mailboxes + inboxes from this rank's slab alone ---------------------------------------
// The slab arrives with GLOBAL column indices; nothing here (or in the callers) ever holds the global matrix.  Index sets are the
// ones the oracle defines: halo = ascending distinct remote global columns (hence grouped by owner, ascending);
// local column = c - lo, remote column = n_local + rank in the halo list; send list for peer q = the local rows q's halo names,
// ascending.  The exchange of counts, index lists and CUDA IPC handles runs over the NCCL communicator of `d`.
namespace {
__global__ void mark_remote_kernel(int64_t nnz, const int* __restrict__ inds, int lo, int hi, int* mark) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    const int c = inds[p];
    if (c < lo || c >= hi) mark[c] = 1;
}
__global__ void compact_halo_kernel(int64_t n_global, const int* __restrict__ mark, const int* __restrict__ rank, int* halo_cols) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_global && mark[c]) halo_cols[rank[c]] = (int)c;
}
__global__ void renumber_kernel(int64_t nnz, int* inds, int lo, int hi, const int* __restrict__ rank) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    const int c = inds[p];
    inds[p] = (c >= lo && c < hi) ? c - lo : (hi - lo) + rank[c];
}
// off[q] = first halo slot whose column is >= bounds[q]; need_idx[i] = halo column as a row index local to its owner
__global__ void owner_offsets_kernel(int P, const long long* __restrict__ bounds, int n_halo, const int* __restrict__ halo_cols, long long* off) {
    const int q = threadIdx.x;
    if (q > P) return;
    int lo = 0, hi = n_halo;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((long long)halo_cols[mid] < bounds[q]) lo = mid + 1; else hi = mid;
    }
    off[q] = lo;
}
__global__ void need_idx_kernel(int P, const long long* __restrict__ bounds, int n_halo, const int* __restrict__ halo_cols, int* need_idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_halo) return;
    const long long c = halo_cols[i];
    int q = 0;
    while (q + 1 < P && bounds[q + 1] <= c) ++q;
    need_idx[i] = (int)(c - bounds[q]);
}
}  // namespace

namespace mpg { int scan_i32(mpg_ctx* ctx, int64_t n, const int* in, int* out); }

// all-gather `bytes` bytes per rank (host in, host out) through a device staging buffer
static int allgather_host(mpg_ctx* ctx, mpg_dist* d, const void* mine, size_t bytes, void* all) {
    char* stage = nullptr;
    MPG_CUDA(ctx, cudaMalloc(&stage, bytes * (size_t)(d->world + 1)));
    struct Free { char* p; ~Free() { cudaFree(p); } } guard{stage};
    MPG_CUDA(ctx, cudaMemcpyAsync(stage, mine, bytes, cudaMemcpyHostToDevice, ctx->stream));
    MPG_NCCL(ctx, nccl().AllGather(stage, stage + bytes, bytes, ncclChar, d->comm, ctx->stream));
    MPG_CUDA(ctx, cudaMemcpyAsync(all, stage + bytes, bytes * (size_t)d->world, cudaMemcpyDeviceToHost, ctx->stream));
    MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return MPG_OK;
}

// device part of the set-up, no communication: halo list, column renumbering, owner offsets and the row lists to request.
// halo_cols / need_idx are cudaMalloc'ed here (caller frees); off gets P + 1 entries.
static int partition_slab_device(mpg_ctx* ctx, int64_t n_global, int P, const int64_t* bounds, int me, int64_t nnz_local, int* inds, int* nh_out,
                                 int** halo_cols_out, int** need_idx_out, std::vector<long long>& off) {
    const int64_t lo = bounds[me], hi = bounds[me + 1];
    int *mark = nullptr, *rank = nullptr;
    long long *bounds_d = nullptr, *off_d = nullptr;
    struct Tmp { int** a; int** b; long long** e; long long** f; ~Tmp() { cudaFree(*a); cudaFree(*b); cudaFree(*e); cudaFree(*f); } } tmp{&mark, &rank, &bounds_d, &off_d};
    MPG_CUDA(ctx, cudaMalloc(&mark, sizeof(int) * (size_t)(n_global + 1)));
    MPG_CUDA(ctx, cudaMalloc(&rank, sizeof(int) * (size_t)(n_global + 1)));
    MPG_CUDA(ctx, cudaMalloc(&bounds_d, sizeof(long long) * (size_t)(P + 1)));
    MPG_CUDA(ctx, cudaMalloc(&off_d, sizeof(long long) * (size_t)(P + 1)));
    MPG_CUDA(ctx, cudaMemsetAsync(mark, 0, sizeof(int) * (size_t)(n_global + 1), ctx->stream));
    std::vector<long long> bl((size_t)P + 1);
    for (int q = 0; q <= P; ++q) bl[(size_t)q] = bounds[q];
    MPG_CUDA(ctx, cudaMemcpyAsync(bounds_d, bl.data(), sizeof(long long) * bl.size(), cudaMemcpyHostToDevice, ctx->stream));
    if (nnz_local > 0) {
        mark_remote_kernel<<<(int)cdiv(nnz_local, 256), 256, 0, ctx->stream>>>(nnz_local, inds, (int)lo, (int)hi, mark);
        MPG_CHECK_LAUNCH(ctx);
    }
    MPG_TRY(mpg::scan_i32(ctx, n_global, mark, rank));
    int nh = 0;
    MPG_CUDA(ctx, cudaMemcpyAsync(&nh, rank + n_global, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    MPG_CUDA(ctx, cudaMalloc(halo_cols_out, sizeof(int) * (size_t)std::max(nh, 1)));
    MPG_CUDA(ctx, cudaMalloc(need_idx_out, sizeof(int) * (size_t)std::max(nh, 1)));
    if (n_global > 0) {
        compact_halo_kernel<<<(int)cdiv(n_global, 256), 256, 0, ctx->stream>>>(n_global, mark, rank, *halo_cols_out);
        MPG_CHECK_LAUNCH(ctx);
    }
    if (nnz_local > 0) {
        renumber_kernel<<<(int)cdiv(nnz_local, 256), 256, 0, ctx->stream>>>(nnz_local, inds, (int)lo, (int)hi, rank);
        MPG_CHECK_LAUNCH(ctx);
    }
    owner_offsets_kernel<<<1, 32, 0, ctx->stream>>>(P, bounds_d, nh, *halo_cols_out, off_d);
    MPG_CHECK_LAUNCH(ctx);
    if (nh > 0) {
        need_idx_kernel<<<(int)cdiv(nh, 256), 256, 0, ctx->stream>>>(P, bounds_d, nh, *halo_cols_out, *need_idx_out);
        MPG_CHECK_LAUNCH(ctx);
    }
    off.assign((size_t)P + 1, 0);
    MPG_CUDA(ctx, cudaMemcpyAsync(off.data(), off_d, sizeof(long long) * off.size(), cudaMemcpyDeviceToHost, ctx->stream));
    MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *nh_out = nh;
    return MPG_OK;
}

// The device kernels of the set-up alone (no communicator): what rank `rank` of `P` would compute for its slab.  Lets a single GPU
// check every rank's halo list, renumbering and request lists against the oracle (tests/test_ops_gpu.py).
extern "C" int mpg_partition_slab_dev(mpg_ctx* ctx, int64_t n_global, int P, const int64_t* bounds, int rank, int64_t nnz_local, int* inds,
                                      int64_t* n_halo, int* halo_cols_dev, int* need_idx_dev, int64_t cap, int64_t* owner_off_host) {
    MPG_REQUIRE(ctx, bounds && P >= 1 && rank >= 0 && rank < P && (inds || nnz_local == 0) && n_halo && n_global >= 0 && n_global < (int64_t)2147483647,
                "partition_slab_dev: bad argument");
    int nh = 0;
    int *hc = nullptr, *ni = nullptr;
    std::vector<long long> off;
    const int rc = partition_slab_device(ctx, n_global, P, bounds, rank, nnz_local, inds, &nh, &hc, &ni, off);
    struct Free { int* a; int* b; ~Free() { cudaFree(a); cudaFree(b); } } guard{hc, ni};
    if (rc != MPG_OK) return rc;
    *n_halo = nh;
    if (nh > cap && (halo_cols_dev || need_idx_dev)) return fail(ctx, MPG_ERR_ARG, "partition_slab_dev: output capacity too small");
    if (halo_cols_dev && nh) MPG_CUDA(ctx, cudaMemcpyAsync(halo_cols_dev, hc, sizeof(int) * (size_t)nh, cudaMemcpyDeviceToDevice, ctx->stream));
    if (need_idx_dev && nh) MPG_CUDA(ctx, cudaMemcpyAsync(need_idx_dev, ni, sizeof(int) * (size_t)nh, cudaMemcpyDeviceToDevice, ctx->stream));
    MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (owner_off_host) for (int q = 0; q <= P; ++q) owner_off_host[q] = off[(size_t)q];
    return MPG_OK;
}

extern "C" int mpg_dist_setup(mpg_ctx* ctx, mpg_dist* d, int64_t n_global, const int64_t* bounds, int64_t nnz_local, int* inds, int64_t* n_halo_out) {
    MPG_REQUIRE(ctx, d && bounds && (inds || nnz_local == 0) && n_global >= 0 && n_global < (int64_t)2147483647 && nnz_local >= 0, "dist_setup: bad argument");
    MPG_REQUIRE(ctx, d->world <= kMaxPeers, "dist_setup: at most 8 ranks");
    const int P = d->world, me = d->rank;
    for (int q = 0; q < P; ++q) MPG_REQUIRE(ctx, bounds[q] <= bounds[q + 1], "dist_setup: bounds must be non-decreasing");
    MPG_REQUIRE(ctx, bounds[0] == 0 && bounds[P] == n_global, "dist_setup: bounds must cover [0, n_global]");
    const int64_t n_local = bounds[me + 1] - bounds[me];
    MPG_REQUIRE(ctx, n_local >= 1, "dist_setup: every rank must own at least one row (a rank without rows would never join the reductions)");
    int nh = 0;
    int* need_idx = nullptr;
    std::vector<long long> off;
    cudaFree(d->halo_cols_dev); d->halo_cols_dev = nullptr;
    MPG_TRY(partition_slab_device(ctx, n_global, P, bounds, me, nnz_local, inds, &nh, &d->halo_cols_dev, &need_idx, off));
    struct Free { int* a; ~Free() { cudaFree(a); } } guard{need_idx};
    // counts: need[q] = halo slots owned by q; all-gathered into cnt[r][q]
    std::vector<long long> need((size_t)P), cnt((size_t)P * P);
    for (int q = 0; q < P; ++q) need[(size_t)q] = off[(size_t)q + 1] - off[(size_t)q];
    MPG_REQUIRE(ctx, need[(size_t)me] == 0, "dist_setup: internal error (own columns in the halo)");
    MPG_TRY(allgather_host(ctx, d, need.data(), sizeof(long long) * (size_t)P, cnt.data()));
    std::vector<long long> soff((size_t)P + 1, 0);
    for (int q = 0; q < P; ++q) soff[(size_t)q + 1] = soff[(size_t)q] + cnt[(size_t)q * P + me];   // what q needs from me
    cudaFree(d->send_idx_own); d->send_idx_own = nullptr;
    MPG_CUDA(ctx, cudaMalloc(&d->send_idx_own, sizeof(int) * (size_t)std::max<long long>(soff[(size_t)P], 1)));
    MPG_NCCL(ctx, nccl().GroupStart());
    for (int q = 0; q < P; ++q) {
        if (q == me) continue;
        if (need[(size_t)q] > 0) MPG_NCCL(ctx, nccl().Send(need_idx + off[(size_t)q], (size_t)need[(size_t)q], ncclInt32, q, d->comm, ctx->stream));
        const long long sc = cnt[(size_t)q * P + me];
        if (sc > 0) MPG_NCCL(ctx, nccl().Recv(d->send_idx_own + soff[(size_t)q], (size_t)sc, ncclInt32, q, d->comm, ctx->stream));
    }
    MPG_NCCL(ctx, nccl().GroupEnd());
    MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // the plan (same call the host-built path makes)
    std::vector<int> pr; std::vector<int64_t> scount, roff, rcount; std::vector<const int*> sptr;
    for (int q = 0; q < P; ++q) {
        if (q == me) continue;
        const long long sc = cnt[(size_t)q * P + me], rc = need[(size_t)q];
        if (sc == 0 && rc == 0) continue;
        pr.push_back(q); scount.push_back(sc); sptr.push_back(d->send_idx_own + soff[(size_t)q]); roff.push_back(off[(size_t)q]); rcount.push_back(rc);
    }
    MPG_TRY(mpg_dist_set_partition(ctx, d, n_global, n_local, nh, (int)pr.size(), pr.data(), scount.data(), sptr.data(), roff.data(), rcount.data()));
    // mailboxes and inboxes: IPC handles (and where each rank's rows land in the others' halos) all-gathered over NCCL
    struct Msg { unsigned char mbox[64]; unsigned char inbox[64]; long long n_halo; long long recv_off[kMaxPeers]; };
    Msg mine;
    memset(&mine, 0, sizeof(mine));
    MPG_TRY(mpg_dist_mailbox_handle(ctx, d, mine.mbox));
    MPG_TRY(mpg_dist_halo_handle(ctx, d, mine.inbox));
    mine.n_halo = nh;
    for (int q = 0; q < P; ++q) mine.recv_off[q] = off[(size_t)q];
    std::vector<Msg> all((size_t)P);
    MPG_TRY(allgather_host(ctx, d, &mine, sizeof(Msg), all.data()));
    std::vector<unsigned char> hm((size_t)P * 64), hi_((size_t)P * 64);
    for (int q = 0; q < P; ++q) { memcpy(hm.data() + (size_t)q * 64, all[(size_t)q].mbox, 64); memcpy(hi_.data() + (size_t)q * 64, all[(size_t)q].inbox, 64); }
    MPG_TRY(mpg_dist_open_mailboxes(ctx, d, hm.data()));
    std::vector<int64_t> remote_off(std::max<size_t>(pr.size(), 1)), remote_nh(std::max<size_t>(pr.size(), 1));
    for (size_t i = 0; i < pr.size(); ++i) { remote_off[i] = all[(size_t)pr[i]].recv_off[me]; remote_nh[i] = all[(size_t)pr[i]].n_halo; }
    MPG_TRY(mpg_dist_open_halo(ctx, d, hi_.data(), remote_off.data(), remote_nh.data()));
    // nobody may start pushing before everybody has mapped everybody: one more collective as a barrier
    long long one = 1;
    std::vector<long long> ones((size_t)P);
    MPG_TRY(allgather_host(ctx, d, &one, sizeof(long long), ones.data()));
    if (n_halo_out) *n_halo_out = nh;
    return MPG_OK;
}

// global column of every halo slot, ascending (the bit-exact artefact of SURVEY.md §8e); host array of n_halo int64
extern "C" int mpg_dist_halo_cols(mpg_ctx* ctx, const mpg_dist* d, int64_t* halo_cols_host) {
    MPG_REQUIRE(ctx, d && halo_cols_host && (d->halo_cols_dev || d->n_halo == 0), "dist_halo_cols: no native plan (use mpg_dist_setup)");
    std::vector<int> h((size_t)d->n_halo);
    if (d->n_halo) MPG_CUDA(ctx, cudaMemcpyAsync(h.data(), d->halo_cols_dev, sizeof(int) * h.size(), cudaMemcpyDeviceToHost, ctx->stream));
    MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < h.size(); ++i) halo_cols_host[i] = h[i];
    return MPG_OK;
}
// send list for the i-th neighbour of the plan (local row indices, ascending): sizes, then contents
extern "C" int mpg_dist_peer_info(mpg_ctx* ctx, const mpg_dist* d, int i, int* peer_rank, int64_t* send_count, int64_t* recv_offset, int64_t* recv_count, int* send_idx_host) {
    MPG_REQUIRE(ctx, d && i >= 0, "dist_peer_info: bad argument");
    if (i >= (int)d->peers.size()) return MPG_ERR_ARG;
    const auto& p = d->peers[(size_t)i];
    if (peer_rank) *peer_rank = p.rank;
    if (send_count) *send_count = p.send_count;
    if (recv_offset) *recv_offset = p.recv_offset;
    if (recv_count) *recv_count = p.recv_count;
    if (send_idx_host && p.send_count) {
        MPG_CUDA(ctx, cudaMemcpyAsync(send_idx_host, p.send_idx, sizeof(int) * (size_t)p.send_count, cudaMemcpyDeviceToHost, ctx->stream));
        MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return MPG_OK;
}

extern "C" int mpg_ctx_attach_dist(mpg_ctx* ctx, mpg_dist* d) {
    if (!ctx) return MPG_ERR_ARG;
    ctx->dist = (d && d->world > 1) ? d : nullptr;   // a single-rank communicator behaves exactly like no communicator
    return MPG_OK;
}

extern "C" int mpg_dist_info(const mpg_dist* d, int* rank, int* world, int64_t* n_global, int64_t* n_local, int64_t* n_halo) {
    if (!d) return MPG_ERR_ARG;
    if (rank) *rank = d->rank;
    if (world) *world = d->world;
    if (n_global) *n_global = d->n_global;
    if (n_local) *n_local = d->n_local;
    if (n_halo) *n_halo = d->n_halo;
    return MPG_OK;
}

namespace mpg {

Epi make_epi(mpg_ctx* ctx, int kind, void* p0, void* p1, double alpha, double beta) {
    Epi e{kind, p0, p1, alpha, beta, nullptr, PeerComm()};
    mpg_dist* d = ctx->dist;
    if (!d) return e;
    if (d->peer_ready && ctx->tune.dist_peer_reduce) {
        e.peer.world = d->world;
        e.peer.rank = d->rank;
        e.peer.seq = ++d->seq;   // every rank issues the same sequence of reductions
        e.peer.err = ctx->dev_err_d;
        e.peer.spin_limit_ns = (unsigned long long)std::max(ctx->tune.spin_limit_ms, 0) * 1000000ull;
        for (int q = 0; q < d->world; ++q) {
            e.peer.mbox[q] = static_cast<double*>(d->mbox_map[q]);
            e.peer.flag[q] = reinterpret_cast<unsigned long long*>(static_cast<char*>(d->mbox_map[q]) + d->mbox_data_bytes());
            e.peer.ll[q] = reinterpret_cast<ulonglong2*>(static_cast<char*>(d->mbox_map[q]) + d->mbox_data_bytes() + d->mbox_flag_bytes());
        }
        e.peer.use_ll = ctx->tune.dist_ll_reduce;
    } else {
        e.raw = ctx->red_raw;
    }
    return e;
}

int dist_finish_reduction(mpg_ctx* ctx, const Epi& e, int count, int tbytes) {
    if (!ctx->dist || !e.raw || count <= 0) return MPG_OK;
    mpg_dist* d = ctx->dist;
    ProfScope prof(ctx, MPG_PROF_SMALL, 0.0);
    const ncclRedOp_t op = (e.kind == EPI_MAX) ? ncclMax : ncclSum;
    MPG_NCCL(ctx, nccl().AllReduce(e.raw, e.raw, (size_t)count, ncclDouble, op, d->comm, ctx->stream));
    const int grid = (count + 127) / 128;
    if (tbytes == 4) epilogue_kernel<float><<<grid, 128, 0, ctx->stream>>>(e, count);
    else epilogue_kernel<double><<<grid, 128, 0, ctx->stream>>>(e, count);
    MPG_CHECK_LAUNCH(ctx);
    return MPG_OK;
}

// x_ext: n_local owned values followed by n_halo halo slots.
// halo_begin sends this rank's rows to the neighbours; halo_finish makes the neighbours' rows visible in the halo
// tail.  Work that does not read the tail (the SpMV tiles without halo columns) goes between the two.
template <class T>
int halo_begin(mpg_ctx* ctx, T* x_ext) {
    mpg_dist* d = ctx->dist;
    if (!d || d->peers.empty()) return MPG_OK;
    ProfScope prof(ctx, MPG_PROF_SMALL, 0.0);
    if (d->halo_ready && ctx->tune.dist_peer_halo) {
        // push over peer memory: gather-and-store kernel, kPushBlocksPerPeer blocks per neighbour
        const unsigned long long seq = ++d->halo_seq;
        const int slot = (int)(seq & 1);
        PushArgs pa;
        pa.npeers = (int)d->peers.size();
        pa.seq = seq;
        pa.counters = d->push_counters;
        for (size_t i = 0; i < d->peers.size(); ++i) {
            const auto& p = d->peers[i];
            char* inbox = static_cast<char*>(d->inbox_map[p.rank]);
            const size_t their_data = 2 * (size_t)std::max<int64_t>(d->remote_nhalo[i], 1) * 8;
            pa.send_idx[i] = p.send_idx;
            pa.count[i] = p.send_count;
            pa.dst[i] = inbox + (size_t)slot * (size_t)std::max<int64_t>(d->remote_nhalo[i], 1) * 8 + (size_t)d->remote_off[i] * sizeof(T);
            pa.flag[i] = reinterpret_cast<unsigned long long*>(inbox + their_data) + slot * kMaxPeers + d->rank;
        }
        pa.bpp = push_blocks(d->max_send_count);
        halo_push_kernel<T><<<pa.npeers * pa.bpp, 256, 0, ctx->stream>>>(pa, x_ext);
        MPG_CHECK_LAUNCH(ctx);
        return MPG_OK;
    }
    T* sbuf = static_cast<T*>(d->send_buf);
    for (const auto& p : d->peers) {
        if (p.send_count == 0) continue;
        pack_kernel<T><<<(int)cdiv(p.send_count, 256), 256, 0, ctx->stream>>>(p.send_count, p.send_idx, x_ext, sbuf + p.send_offset);
        MPG_CHECK_LAUNCH(ctx);
    }
    const ncclDataType_t dt = sizeof(T) == 4 ? ncclFloat : ncclDouble;
    MPG_NCCL(ctx, nccl().GroupStart());
    for (const auto& p : d->peers) {
        if (p.send_count > 0) MPG_NCCL(ctx, nccl().Send(sbuf + p.send_offset, (size_t)p.send_count, dt, p.rank, d->comm, ctx->stream));
        if (p.recv_count > 0) MPG_NCCL(ctx, nccl().Recv(x_ext + d->n_local + p.recv_offset, (size_t)p.recv_count, dt, p.rank, d->comm, ctx->stream));
    }
    MPG_NCCL(ctx, nccl().GroupEnd());
    return MPG_OK;
}

template <class T>
int halo_finish(mpg_ctx* ctx, T* x_ext) {
    mpg_dist* d = ctx->dist;
    if (!d || d->peers.empty()) return MPG_OK;
    if (!(d->halo_ready && ctx->tune.dist_peer_halo)) return MPG_OK;   // NCCL path: the receive was enqueued by halo_begin
    ProfScope prof(ctx, MPG_PROF_SMALL, 0.0);
    // wait-and-move kernel: spin on the neighbours' exchange numbers, then inbox -> halo tail
    const unsigned long long seq = d->halo_seq;
    const int slot = (int)(seq & 1);
    WaitArgs wa;
    wa.npeers = (int)d->peers.size();
    wa.err = ctx->dev_err_d;
    wa.spin_limit_ns = (unsigned long long)std::max(ctx->tune.spin_limit_ms, 0) * 1000000ull;
    for (size_t i = 0; i < d->peers.size(); ++i)
        wa.flag[i] = reinterpret_cast<const unsigned long long*>(static_cast<char*>(d->inbox_own) + d->inbox_data_bytes()) + slot * kMaxPeers + d->peers[i].rank;
    const T* inbox = reinterpret_cast<const T*>(static_cast<char*>(d->inbox_own) + (size_t)slot * (size_t)std::max<int64_t>(d->n_halo, 1) * 8);
    const int grid = (int)std::min<int64_t>(32, std::max<int64_t>(1, cdiv(d->n_halo, 256 * 8)));
    halo_wait_copy_kernel<T><<<grid, 256, 0, ctx->stream>>>(wa, seq, inbox, x_ext + d->n_local, (long long)d->n_halo);
    MPG_CHECK_LAUNCH(ctx);
    return MPG_OK;
}

// ---- fused halo: push straight into the neighbours' basis columns ---------------------------------------------------------
// Collective, once per solve: every rank publishes the CUDA IPC handle of its basis allocation with its leading dimension and row
// count; neighbours are (re)mapped only when their handle changed.  Returns with d->basis_ready = all neighbours mapped.
int dist_exchange_basis(mpg_ctx* ctx, void* V, int64_t ldv, int tsize) {
    mpg_dist* d = ctx->dist;
    if (!d) return MPG_OK;
    d->basis_ready = false;
    if (!(d->halo_ready && ctx->tune.dist_peer_halo && ctx->tune.dist_fuse_halo) || d->world > kMaxPeers) return MPG_OK;
    struct Msg { unsigned char handle[64]; long long ldv, n_local, tsize, pad; };
    static_assert(sizeof(Msg) == 96, "hand-shake message");
    Msg mine;
    memset(&mine, 0, sizeof(mine));
    cudaIpcMemHandle_t h;
    MPG_CUDA(ctx, cudaIpcGetMemHandle(&h, V));
    memcpy(mine.handle, &h, 64);
    mine.ldv = ldv; mine.n_local = d->n_local; mine.tsize = tsize;
    if (!d->xchg_stage) MPG_CUDA(ctx, cudaMalloc(&d->xchg_stage, sizeof(Msg) * (size_t)(kMaxPeers + 1)));
    std::vector<Msg> all((size_t)d->world);
    MPG_CUDA(ctx, cudaMemcpyAsync(d->xchg_stage, &mine, sizeof(Msg), cudaMemcpyHostToDevice, ctx->stream));
    MPG_NCCL(ctx, nccl().AllGather(d->xchg_stage, d->xchg_stage + sizeof(Msg), sizeof(Msg), ncclChar, d->comm, ctx->stream));
    MPG_CUDA(ctx, cudaMemcpyAsync(all.data(), d->xchg_stage + sizeof(Msg), sizeof(Msg) * all.size(), cudaMemcpyDeviceToHost, ctx->stream));
    MPG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (const auto& p : d->peers) {
        const Msg& m = all[(size_t)p.rank];
        auto& vp = d->vpeer[p.rank];
        if (!vp.valid || memcmp(vp.handle, m.handle, 64) != 0) {
            if (vp.base) { cudaIpcCloseMemHandle(vp.base); vp.base = nullptr; }
            cudaIpcMemHandle_t ph;
            memcpy(&ph, m.handle, 64);
            MPG_CUDA(ctx, cudaIpcOpenMemHandle(&vp.base, ph, cudaIpcMemLazyEnablePeerAccess));
            memcpy(vp.handle, m.handle, 64);
            vp.valid = true;
        }
        vp.ldv = m.ldv; vp.n_local = m.n_local; vp.tsize = (int)m.tsize;
        if (vp.tsize != tsize) return fail(ctx, MPG_ERR_STATE, "dist: ranks solve in different precisions");
    }
    d->basis_ready = true;
    return MPG_OK;
}
bool dist_basis_ready(mpg_ctx* ctx) { return ctx->dist && ctx->dist->basis_ready; }

// push descriptor for basis column `col`: our boundary rows go to rows [n_q + off, ...) of column `col` of every neighbour q
template <class T>
int halo_direct_args(mpg_ctx* ctx, int64_t col, PushArgs* pa) {
    mpg_dist* d = ctx->dist;
    const unsigned long long seq = ++d->halo_seq;
    const int slot = (int)(seq & 1);
    pa->npeers = (int)d->peers.size();
    pa->seq = seq;
    pa->counters = d->push_counters;
    for (size_t i = 0; i < d->peers.size(); ++i) {
        const auto& p = d->peers[i];
        const auto& vp = d->vpeer[p.rank];
        char* inbox = static_cast<char*>(d->inbox_map[p.rank]);
        const size_t their_data = 2 * (size_t)std::max<int64_t>(d->remote_nhalo[i], 1) * 8;
        pa->send_idx[i] = p.send_idx;
        pa->count[i] = p.send_count;
        pa->dst[i] = static_cast<T*>(vp.base) + (size_t)col * (size_t)vp.ldv + (size_t)vp.n_local + (size_t)d->remote_off[i];
        pa->flag[i] = reinterpret_cast<unsigned long long*>(inbox + their_data) + slot * kMaxPeers + d->rank;
    }
    pa->bpp = push_blocks(d->max_send_count);
    return MPG_OK;
}
template int halo_direct_args<float>(mpg_ctx*, int64_t, PushArgs*);
template int halo_direct_args<double>(mpg_ctx*, int64_t, PushArgs*);

// stand-alone push of the column that starts at `x` (first vector of a cycle, unfused tail)
template <class T>
int halo_push_direct(mpg_ctx* ctx, const T* x, int64_t col) {
    mpg_dist* d = ctx->dist;
    if (!d || d->peers.empty()) return MPG_OK;
    ProfScope prof(ctx, MPG_PROF_SMALL, 0.0);
    PushArgs pa;
    MPG_TRY(halo_direct_args<T>(ctx, col, &pa));
    halo_push_kernel<T><<<pa.npeers * pa.bpp, 256, 0, ctx->stream>>>(pa, x);
    MPG_CHECK_LAUNCH(ctx);
    return MPG_OK;
}
template int halo_push_direct<float>(mpg_ctx*, const float*, int64_t);
template int halo_push_direct<double>(mpg_ctx*, const double*, int64_t);

// what the consumer of the latest exchange waits for
int halo_wait_args(mpg_ctx* ctx, HaloWait* hw) {
    mpg_dist* d = ctx->dist;
    hw->npeers = 0;
    if (!d || d->peers.empty()) return MPG_OK;
    const unsigned long long seq = d->halo_seq;
    const int slot = (int)(seq & 1);
    hw->npeers = (int)d->peers.size();
    hw->seq = seq;
    hw->err = ctx->dev_err_d;
    hw->spin_limit_ns = (unsigned long long)std::max(ctx->tune.spin_limit_ms, 0) * 1000000ull;
    for (size_t i = 0; i < d->peers.size(); ++i)
        hw->flag[i] = reinterpret_cast<const unsigned long long*>(static_cast<char*>(d->inbox_own) + d->inbox_data_bytes()) + slot * kMaxPeers + d->peers[i].rank;
    return MPG_OK;
}
int halo_wait_only(mpg_ctx* ctx) {
    HaloWait hw;
    MPG_TRY(halo_wait_args(ctx, &hw));
    if (hw.npeers == 0) return MPG_OK;
    ProfScope prof(ctx, MPG_PROF_SMALL, 0.0);
    halo_wait_kernel<<<1, 32, 0, ctx->stream>>>(hw);
    MPG_CHECK_LAUNCH(ctx);
    return MPG_OK;
}

template <class T>
int halo_exchange(mpg_ctx* ctx, T* x_ext) {
    MPG_TRY(halo_begin<T>(ctx, x_ext));
    return halo_finish<T>(ctx, x_ext);
}
template int halo_begin<float>(mpg_ctx*, float*);
template int halo_begin<double>(mpg_ctx*, double*);
template int halo_finish<float>(mpg_ctx*, float*);
template int halo_finish<double>(mpg_ctx*, double*);
template int halo_exchange<float>(mpg_ctx*, float*);
template int halo_exchange<double>(mpg_ctx*, double*);

int64_t dist_halo(mpg_ctx* ctx) { return ctx->dist ? ctx->dist->n_halo : 0; }
int64_t dist_nlocal(mpg_ctx* ctx) { return ctx->dist ? ctx->dist->n_local : -1; }
int64_t dist_nglobal(mpg_ctx* ctx) { return ctx->dist ? ctx->dist->n_global : -1; }
int dist_world(mpg_ctx* ctx) { return ctx->dist ? ctx->dist->world : 1; }

}  // namespace mpg

extern "C" int mpg_halo_exchange_f32(mpg_ctx* ctx, float* x_ext) { return mpg::halo_exchange<float>(ctx, x_ext); }
extern "C" int mpg_halo_exchange_f64(mpg_ctx* ctx, double* x_ext) { return mpg::halo_exchange<double>(ctx, x_ext); }
extern "C" int mpg_allreduce_sum_f64(mpg_ctx* ctx, double* buf, int64_t count) {
    if (!ctx->dist) return MPG_OK;
    MPG_NCCL(ctx, nccl().AllReduce(buf, buf, (size_t)count, ncclDouble, ncclSum, ctx->dist->comm, ctx->stream));
    return MPG_OK;
}
