"""world_size-2 (and 3) gloo tests of the multi-GPU host logic on CPU: the partition / halo plan built by
icl-mixed-precision-gmres_b200/dist.py (the same tensor code that runs on the GPUs) must reproduce the oracle's index
sets bit for bit, and a distributed SpMV + dot emulated with the plan (pack -> exchange -> local slab product,
all-reduced partial sums) must equal the global result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, spec, out_q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import gmres_b200 as g
        import oracle as orc
        import scipy.sparse as sp
        rm, ind, val = orc.gen(spec)
        n = len(rm) - 1
        part = g.dist.build_partition(torch.from_numpy(rm), torch.from_numpy(ind), torch.from_numpy(val), n, rank, world)
        # --- index sets: bit-exact vs the oracle (and vs the library's host helper, tested in test_abi_cpu.py) ---
        halo_o, li_o = orc.partition_local(n, world, rank, rm, ind)
        assert np.array_equal(part.halo_cols.numpy(), halo_o), "halo_cols differ"
        assert np.array_equal(part.inds.numpy(), li_o), "local column indices differ"
        b = orc.partition_bounds(n, world)
        assert (part.lo, part.hi) == (b[rank], b[rank + 1])
        assert np.array_equal(part.row_map.numpy(), rm[part.lo:part.hi + 1] - rm[part.lo])
        # recv ranges tile the halo in owner order; send lists are ascending local rows of this rank
        off = 0
        for p in part.peers:
            assert p["recv_offset"] == off
            seg = part.halo_cols[off:off + p["recv_count"]].numpy()
            assert np.all((seg >= b[p["rank"]]) & (seg < b[p["rank"] + 1]))
            off += p["recv_count"]
            s = p["send_idx"].numpy()
            assert np.all(np.diff(s) > 0) and (len(s) == 0 or (s.min() >= 0 and s.max() < part.n_local))
        assert off == part.n_halo
        # --- distributed SpMV with the plan (pack -> exchange -> local product) ---
        x = np.random.default_rng(5).standard_normal(n)
        x_ext = np.concatenate([x[part.lo:part.hi], np.zeros(part.n_halo)])
        sends = {p["rank"]: x_ext[:part.n_local][p["send_idx"].numpy()] for p in part.peers}
        gathered = [None] * world
        dist.all_gather_object(gathered, sends)
        for p in part.peers:
            got = gathered[p["rank"]][rank]
            assert len(got) == p["recv_count"]
            x_ext[part.n_local + p["recv_offset"]:part.n_local + p["recv_offset"] + p["recv_count"]] = got
        assert np.array_equal(x_ext[part.n_local:], x[part.halo_cols.numpy()])   # the halo holds exactly the remote entries
        A_loc = sp.csr_matrix((part.vals.numpy(), part.inds.numpy(), part.row_map.numpy()), shape=(part.n_local, part.n_local + part.n_halo))
        y_loc = A_loc @ x_ext
        y_ref = (sp.csr_matrix((val, ind, rm), shape=(n, n)) @ x)[part.lo:part.hi]
        assert np.allclose(y_loc, y_ref, rtol=1e-13, atol=1e-13)
        # --- all-reduced dot / gemv-T partials == global values ---
        V = np.random.default_rng(7).standard_normal((n, 5))
        t = torch.from_numpy(V[part.lo:part.hi].T @ x[part.lo:part.hi])
        dist.all_reduce(t)
        assert np.allclose(t.numpy(), V.T @ x, rtol=1e-12)
        out_q.put((rank, "ok", part.n_local, part.n_halo, len(part.peers)))
    except Exception as e:  # noqa: BLE001
        import traceback
        out_q.put((rank, "fail: " + repr(e) + "\n" + traceback.format_exc(), 0, 0, 0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,spec", [(2, "lap2d:12"), (2, "cd27:6"), (2, "powerlaw:500"), (3, "cd27:6")])
def test_partition_and_halo_plan_gloo(world, spec):
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = 29500 + (os.getpid() % 1000) + world * 7 + len(spec)
    procs = [ctxm.Process(target=_worker, args=(r, world, port, spec, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for r in sorted(res):
        assert r[1] == "ok", r[1]
    if spec.startswith("cd27"):
        N = int(spec.split(":")[1])
        # z-slab partition of the 27-point stencil: each rank's halo is one N^2 plane per neighbouring slab
        for r in sorted(res):
            nb = (1 if r[0] > 0 else 0) + (1 if r[0] < world - 1 else 0)
            assert r[4] == nb and r[3] == nb * N * N
