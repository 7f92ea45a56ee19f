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
    MPG_SYM(GetUniqueId); MPG_SYM(CommInitRank); MPG_SYM(CommDestroy); MPG_SYM(AllReduce); MPG_SYM(Send); MPG_SYM(Recv);
    MPG_SYM(GroupStart); MPG_SYM(GroupEnd); MPG_SYM(GetErrorString);
#undef MPG_SYM
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.Send && api.Recv && api.GroupStart && api.GroupEnd;
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
struct PushArgs {
    int npeers;
    const int* send_idx[kMaxPeers];
    long long count[kMaxPeers];
    void* dst[kMaxPeers];                   // neighbour's inbox + slot offset + where our rows go (bytes resolved per type)
    unsigned long long* flag[kMaxPeers];    // neighbour's flag for (slot, this rank)
};
constexpr int kPushBlocksPerPeer = 32;   // one SM cannot keep an NVLink busy with stores; spread every neighbour's rows
template <class T>
__global__ void __launch_bounds__(256) halo_push_kernel(PushArgs a, const T* __restrict__ x, unsigned long long seq, unsigned int* counters) {
    const int q = blockIdx.x / kPushBlocksPerPeer, part = blockIdx.x % kPushBlocksPerPeer;
    T* dst = static_cast<T*>(a.dst[q]);
    const int* idx = a.send_idx[q];
    for (long long i = (long long)part * blockDim.x + threadIdx.x; i < a.count[q]; i += (long long)kPushBlocksPerPeer * blockDim.x) dst[i] = x[idx[i]];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        // the last block of this neighbour publishes the exchange number (its fence + the counter chain order all stores)
        if (atomicAdd(counters + q, 1u) == kPushBlocksPerPeer - 1) {
            counters[q] = 0u;
            __threadfence_system();
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.flag[q]), "l"(seq) : "memory");
        }
    }
}
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
    void* send_buf = nullptr;              // send_total doubles
    // peer-memory mailboxes for the in-kernel all-reduce (common.cuh PeerComm)
    void* mbox_own = nullptr;              // [kMboxSlots][world][kMboxStride] doubles, then [kMboxSlots][world] u64 flags
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
    size_t inbox_data_bytes() const { return 2 * (size_t)std::max<int64_t>(n_halo, 1) * 8; }
    size_t inbox_bytes() const { return inbox_data_bytes() + sizeof(unsigned long long) * 2 * kMaxPeers; }
    size_t mbox_data_bytes() const { return sizeof(double) * (size_t)kMboxSlots * world * kMboxStride; }
    size_t mbox_bytes() const { return mbox_data_bytes() + sizeof(unsigned long long) * (size_t)kMboxSlots * world; }
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
    cudaFree(d->inbox_own);
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
    d->n_global = n_global; d->n_local = n_local; d->n_halo = n_halo;
    d->peers.clear();
    int64_t off = 0;
    for (int i = 0; i < npeers; ++i) {
        MPG_REQUIRE(ctx, peer_ranks[i] >= 0 && peer_ranks[i] < d->world && peer_ranks[i] != d->rank, "dist_set_partition: bad peer rank");
        MPG_REQUIRE(ctx, recv_offsets[i] >= 0 && recv_offsets[i] + recv_counts[i] <= n_halo, "dist_set_partition: halo range out of bounds");
        d->peers.push_back({peer_ranks[i], send_counts[i], off, recv_counts[i], recv_offsets[i], send_idx_dev[i]});
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
        }
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
        for (size_t i = 0; i < d->peers.size(); ++i) {
            const auto& p = d->peers[i];
            char* inbox = static_cast<char*>(d->inbox_map[p.rank]);
            const size_t their_data = 2 * (size_t)std::max<int64_t>(d->remote_nhalo[i], 1) * 8;
            pa.send_idx[i] = p.send_idx;
            pa.count[i] = p.send_count;
            pa.dst[i] = inbox + (size_t)slot * (size_t)std::max<int64_t>(d->remote_nhalo[i], 1) * 8 + (size_t)d->remote_off[i] * sizeof(T);
            pa.flag[i] = reinterpret_cast<unsigned long long*>(inbox + their_data) + slot * kMaxPeers + d->rank;
        }
        halo_push_kernel<T><<<pa.npeers * kPushBlocksPerPeer, 256, 0, ctx->stream>>>(pa, x_ext, seq, d->push_counters);
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
