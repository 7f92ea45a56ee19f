"""torchrun target: where does the time of one end-to-end step of bench_dist.py go (per piece, synchronised)"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import gmres_b200 as g
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = f"cuda:{local}"
dist.init_process_group("nccl", device_id=torch.device(dev))
ctx = g.Context(local)
if rank == 0 and os.environ.get("MPG_TRACE"):
    ctx.set_tuning("trace", 1)
spec = sys.argv[1] if len(sys.argv) > 1 else "cd27:256"
_, _, _, _, _, n = ctx.gen_params(spec)
bnd = g.dist.bounds(n, world)
lo, hi = bnd[rank], bnd[rank + 1]
rm, ind, val = ctx.gen_slab(spec, lo, hi)
dctx = g.dist.DistContext(ctx, rank, world, native=True)
part = dctx.setup(n, bnd, rm, ind, val)
A = g.dist.local_csr(ctx, part)
b = torch.rand(part.n_local, dtype=torch.float64, device=dev)
val32 = part.vals.float()
dctx.attach()
kw = dict(mode="mixed", orth="cgsr", conv="base", prec="identity", rlen=100, tol=1e-6, max_restarts=1000)
x = torch.zeros(part.n_local, dtype=torch.float64, device=dev)
def T(label, fn, out):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); out[label] = round((time.perf_counter() - t0) * 1e3, 2); return r
res = []
for rep in range(6):
    o = {}
    T("solve_same_A", lambda: ctx.gmres(A, part.vals, b, x.zero_(), vals32=val32, hist_cap=1, **kw), o)
    A2 = T("csr_create", lambda: g.CSR(ctx, part.row_map, part.inds, ncols=part.n_local + part.n_halo), o)
    T("solve_new_A", lambda: ctx.gmres(A2, part.vals, b, x.zero_(), vals32=val32, hist_cap=1, **kw), o)
    T("solve_new_A_again", lambda: ctx.gmres(A2, part.vals, b, x.zero_(), vals32=val32, hist_cap=1, **kw), o)
    T("csr_destroy", lambda: A2.__del__(), o)
    res.append(o)
if rank == 0:
    print(json.dumps(res))
dctx.close(); dist.destroy_process_group()
