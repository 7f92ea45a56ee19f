"""torchrun target: latency of the cross-GPU primitives of the partitioned path, per call (CUDA events, 200 calls):
halo exchange (peer push / NCCL), the in-kernel all-reduce behind nrm2 and a k1-wide gemv-T, local baselines."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import gmres_b200 as g

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
ctx = g.Context(local)
spec = sys.argv[1] if len(sys.argv) > 1 else "cd27:128"
rm, ind, val = ctx.gen(spec)
n = rm.numel() - 1
part = g.dist.build_partition(rm, ind, val, n, rank, world)
del rm, ind, val
nl = part.n_local
x32 = torch.randn(nl + part.n_halo, dtype=torch.float32, device=dev)
small = torch.randn(1024, dtype=torch.float32, device=dev)
out = torch.zeros(8, dtype=torch.float32, device=dev)
k1 = 50
ld = (nl + 31) // 32 * 32
V = torch.randn(ld * k1, dtype=torch.float32, device=dev)
w = torch.randn(ld, dtype=torch.float32, device=dev)
h = torch.zeros(k1 + 2, dtype=torch.float32, device=dev)


def timeit(fn, reps=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t.item()), 2)


res = {"world": world, "spec": spec, "n_local": nl, "n_halo": part.n_halo, "unit": "us per call"}
res["local_nrm2_1k"] = timeit(lambda: ctx.nrm2_dev(small, out))
res["local_nrm2_slab"] = timeit(lambda: ctx.nrm2_dev(x32[:nl], out))
res["local_gemvt_k50"] = timeit(lambda: ctx.gemv(True, nl, k1, 1.0, V, ld, w, 0.0, h))
for peer in (True, False):
    d = g.dist.DistContext(ctx, rank, world, peer_reduce=peer)
    d.set_partition(part)
    d.attach()
    tag = "peer" if peer else "nccl"
    ctx.set_tuning("dist_peer_halo", 1 if peer else 0)
    res[f"halo_exchange_{tag}"] = timeit(lambda: d.halo_exchange(x32))
    res[f"allreduce_nrm2_1k_{tag}"] = timeit(lambda: ctx.nrm2_dev(small, out))
    res[f"allreduce_nrm2_slab_{tag}"] = timeit(lambda: ctx.nrm2_dev(x32[:nl], out))
    res[f"allreduce_gemvt_k50_{tag}"] = timeit(lambda: ctx.gemv(True, nl, k1, 1.0, V, ld, w, 0.0, h))
    d.detach(); d.close()
ctx.set_tuning("dist_peer_halo", 1)
if rank == 0:
    print(json.dumps(res), flush=True)
dist.destroy_process_group()
