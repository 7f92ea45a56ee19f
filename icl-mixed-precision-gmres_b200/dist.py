"""Host-side plumbing of the multi-GPU path (one process per GPU, torch.distributed for rendezvous and set-up only).

build_partition() turns a global CSR into this rank's slab + halo plan with torch tensor ops (any device, so the
world_size-2 gloo tests exercise exactly this code on CPU); the index sets it produces are the bit-exact artefacts
defined by the oracle: rows [floor(r n/P), floor((r+1) n/P)), local columns -> c - lo, remote columns
-> n_local + rank in the ascending list of distinct remote globals.  DistContext hands the plan and an NCCL
communicator to the C library; after attach the solver's reductions are all-reduced and SpMV inputs are halo-filled."""
import ctypes as C

from .binding import Context, CSR, _ptr


def bounds(n, world):
    """equal row blocks: rank r owns rows [floor(r n / P), floor((r + 1) n / P))"""
    return [(r * n) // world for r in range(world + 1)]


def bounds_nnz(row_map_host, world):
    """nnz-balanced split points: bounds[k] = smallest row i with row_map[i] >= floor(k nnz / P) (mpg_partition_bounds_nnz)"""
    import numpy as np
    from .binding import load_library
    rm = np.ascontiguousarray(row_map_host, np.int32)
    out = (C.c_int64 * (world + 1))()
    rc = load_library().mpg_partition_bounds_nnz(C.c_int64(len(rm) - 1), C.c_int(world), rm.ctypes.data_as(C.c_void_p), out)
    if rc != 0:
        raise ValueError("mpg_partition_bounds_nnz failed")
    return list(out)


class Partition:
    """this rank's share of a 1-D row-partitioned CSR matrix"""

    def __init__(self, n_global, rank, world, row_map, inds, vals, halo_cols, peers, b=None):
        self.n_global, self.rank, self.world = n_global, rank, world
        self.row_map, self.inds, self.vals = row_map, inds, vals      # local slab, columns renumbered [local | halo]
        self.halo_cols = halo_cols                                    # int64 global ids, ascending
        self.peers = peers                                            # list of dict(rank, send_idx, recv_offset, recv_count)
        b = bounds(n_global, world) if b is None else list(b)
        self.bounds = b
        self.lo, self.hi = b[rank], b[rank + 1]
        self.n_local, self.n_halo = self.hi - self.lo, int(halo_cols.numel())


def local_slab(row_map, inds, vals, n_global, rank, world, b=None):
    """(row_map_local, local_inds, vals_local, halo_cols) for rank `rank`; pure tensor ops on the inputs' device"""
    import torch
    b = bounds(n_global, world) if b is None else list(b)
    lo, hi = b[rank], b[rank + 1]
    p0, p1 = int(row_map[lo]), int(row_map[hi])
    rm = (row_map[lo:hi + 1] - row_map[lo]).to(torch.int32)
    c = inds[p0:p1].to(torch.int64)
    remote = (c < lo) | (c >= hi)
    halo_cols = torch.unique(c[remote])  # sorted ascending, distinct
    li = torch.where(remote, (hi - lo) + torch.searchsorted(halo_cols, c), c - lo).to(torch.int32)
    return rm.contiguous(), li.contiguous(), vals[p0:p1].clone(), halo_cols   # clone: a slice keeps the parent's (possibly odd) offset


def build_partition(row_map, inds, vals, n_global, rank, world, group=None, b=None):
    """slab + halo plan from the GLOBAL matrix with torch tensor ops (any device): the CPU-testable restatement of the plan
    (tests/test_dist_cpu.py, gloo).  The product path is DistContext.setup (native, from the slab alone).
    Collective: every rank calls it (the send lists come from the peers' halo lists)."""
    import torch
    import torch.distributed as dist
    b = bounds(n_global, world) if b is None else list(b)
    rm, li, v, halo_cols = local_slab(row_map, inds, vals, n_global, rank, world, b)
    owner_lo = torch.tensor(b[:-1], dtype=torch.int64, device=halo_cols.device)
    owner = torch.searchsorted(owner_lo, halo_cols, right=True) - 1 if halo_cols.numel() else halo_cols
    need = {}
    for q in range(world):
        if q == rank or halo_cols.numel() == 0:
            continue
        sel = halo_cols[owner == q]
        if sel.numel():
            need[q] = (sel - b[q]).to(torch.int32).cpu().numpy()   # row indices local to the owner, ascending
    if world > 1:
        all_need = [None] * world
        dist.all_gather_object(all_need, need, group=group)
    else:
        all_need = [need]
    peers = []
    recv_off = 0
    for q in range(world):
        if q == rank:
            continue
        send = all_need[q].get(rank) if all_need[q] else None
        recv_cnt = len(need[q]) if q in need else 0
        if send is None and recv_cnt == 0:
            continue
        send_idx = torch.from_numpy(send).to(halo_cols.device) if send is not None else torch.empty(0, dtype=torch.int32, device=halo_cols.device)
        peers.append(dict(rank=q, send_idx=send_idx, recv_offset=recv_off, recv_count=recv_cnt))
        recv_off += recv_cnt
    assert recv_off == halo_cols.numel()
    return Partition(n_global, rank, world, rm, li, v, halo_cols, peers, b)


class DistContext:
    """NCCL communicator + halo plan attached to a Context (C ABI: mpg_dist_*)."""

    def __init__(self, ctx: Context, rank, world, group=None, peer_reduce=True, native=False):
        import torch
        import torch.distributed as dist
        self.ctx, self.rank, self.world = ctx, rank, world
        self._group = group
        self.h = C.c_void_p()
        idbuf = (C.c_ubyte * 128)()
        if rank == 0:
            ctx._chk(ctx.L.mpg_nccl_unique_id(idbuf))
        payload = [bytes(idbuf)]
        if world > 1:
            dist.broadcast_object_list(payload, src=0, group=group)
        idbuf = (C.c_ubyte * 128).from_buffer_copy(payload[0])
        ctx._chk(ctx.L.mpg_dist_create(ctx.h, idbuf, C.c_int(rank), C.c_int(world), C.byref(self.h)))
        self._keep = None
        # peer-memory mailboxes for the in-kernel all-reduce (CUDA IPC handles all-gathered in rank order)
        self.peer_reduce = False
        if world > 1 and world <= 8 and peer_reduce and not native:   # (native: mpg_dist_setup exchanges the handles itself, over NCCL)
            hb = (C.c_ubyte * 64)()
            ctx._chk(ctx.L.mpg_dist_mailbox_handle(ctx.h, self.h, hb))
            handles = [None] * world
            dist.all_gather_object(handles, bytes(hb), group=group)
            allh = (C.c_ubyte * (64 * world)).from_buffer_copy(b"".join(handles))
            ctx._chk(ctx.L.mpg_dist_open_mailboxes(ctx.h, self.h, allh))
            dist.barrier(group=group)
            self.peer_reduce = True

    def set_partition(self, part: Partition):
        ctx = self.ctx
        n = len(part.peers)
        ranks = (C.c_int * max(n, 1))(*[p["rank"] for p in part.peers])
        scount = (C.c_int64 * max(n, 1))(*[int(p["send_idx"].numel()) for p in part.peers])
        sptr = (C.c_void_p * max(n, 1))(*[p["send_idx"].data_ptr() for p in part.peers])
        roff = (C.c_int64 * max(n, 1))(*[p["recv_offset"] for p in part.peers])
        rcount = (C.c_int64 * max(n, 1))(*[p["recv_count"] for p in part.peers])
        ctx._chk(ctx.L.mpg_dist_set_partition(ctx.h, self.h, C.c_int64(part.n_global), C.c_int64(part.n_local), C.c_int64(part.n_halo), C.c_int(n),
                                              ranks, scount, sptr, roff, rcount))
        self._keep = part  # the send index tensors must outlive the plan
        # peer-memory halo inboxes: exchange IPC handles and, for every neighbour, where our rows land in its halo
        if self.peer_reduce and self.world > 1:
            import torch.distributed as dist
            hb = (C.c_ubyte * 64)()
            ctx._chk(ctx.L.mpg_dist_halo_handle(ctx.h, self.h, hb))
            mine = dict(handle=bytes(hb), n_halo=part.n_halo, recv={p["rank"]: p["recv_offset"] for p in part.peers})
            allinfo = [None] * self.world
            dist.all_gather_object(allinfo, mine, group=self._group)
            allh = (C.c_ubyte * (64 * self.world)).from_buffer_copy(b"".join(i["handle"] for i in allinfo))
            roff = (C.c_int64 * max(n, 1))(*[allinfo[p["rank"]]["recv"].get(self.rank, 0) for p in part.peers])
            rnh = (C.c_int64 * max(n, 1))(*[allinfo[p["rank"]]["n_halo"] for p in part.peers])
            ctx._chk(ctx.L.mpg_dist_open_halo(ctx.h, self.h, allh, roff, rnh))
            dist.barrier(group=self._group)

    def setup(self, n_global, b, row_map_local, inds_global, vals):
        """NATIVE set-up from this rank's slab alone (mpg_dist_setup): inds_global (int32 CUDA tensor, global columns) is renumbered
        in place to [local | halo]; halo / send lists, mailboxes and inboxes are built and exchanged inside the library over NCCL.
        Returns a Partition (halo_cols fetched from the library for tests / diagnostics).  Collective."""
        import numpy as np
        import torch
        ctx = self.ctx
        barr = (C.c_int64 * (self.world + 1))(*[int(v) for v in b])
        nh = C.c_int64()
        ctx._chk(ctx.L.mpg_dist_setup(ctx.h, self.h, C.c_int64(n_global), barr, C.c_int64(inds_global.numel()), _ptr(inds_global), C.byref(nh)))
        self.peer_reduce = self.world > 1
        halo = np.empty(nh.value, np.int64)
        ctx._chk(ctx.L.mpg_dist_halo_cols(ctx.h, self.h, halo.ctypes.data_as(C.c_void_p)))
        peers = []
        i = 0
        while True:
            pr, sc, ro, rc = C.c_int(), C.c_int64(), C.c_int64(), C.c_int64()
            if ctx.L.mpg_dist_peer_info(ctx.h, self.h, C.c_int(i), C.byref(pr), C.byref(sc), C.byref(ro), C.byref(rc), None) != 0:
                break
            sidx = np.empty(sc.value, np.int32)
            ctx._chk(ctx.L.mpg_dist_peer_info(ctx.h, self.h, C.c_int(i), None, None, None, None, sidx.ctypes.data_as(C.c_void_p)))
            peers.append(dict(rank=pr.value, send_idx=torch.from_numpy(sidx), recv_offset=ro.value, recv_count=rc.value))
            i += 1
        part = Partition(n_global, self.rank, self.world, row_map_local, inds_global, vals, torch.from_numpy(halo).to(inds_global.device), peers, b)
        self._keep = part
        return part

    def attach(self):
        self.ctx._chk(self.ctx.L.mpg_ctx_attach_dist(self.ctx.h, self.h))

    def detach(self):
        self.ctx._chk(self.ctx.L.mpg_ctx_attach_dist(self.ctx.h, None))

    def halo_exchange(self, x_ext):
        import torch
        fn = self.ctx.L.mpg_halo_exchange_f32 if x_ext.dtype == torch.float32 else self.ctx.L.mpg_halo_exchange_f64
        self.ctx._chk(fn(self.ctx.h, _ptr(x_ext)))

    def close(self):
        if self.h:
            self.detach()
            self.ctx.L.mpg_dist_destroy(self.h)
            self.h = None


def local_csr(ctx: Context, part: Partition):
    return CSR(ctx, part.row_map, part.inds, ncols=part.n_local + part.n_halo)
