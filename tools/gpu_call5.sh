#!/bin/bash
# round 2, call 5 (2 GPUs): native distributed set-up, fused halo, nnz-balanced split points
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py > gpurun_out/r02e_dist_check_n$N.json 2> gpurun_out/r02e_dist_check_n$N.err; echo "dist_check rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02e_dist_check_n$N.json") if l.startswith("{")][-1])
    print("ok", d["ok"])
    for c in d["cases"]:
        print({k:c[k] for k in ("spec","orth","peer_reduce","split","ok","plan_ok","halo_ok","overlap_ok","replicated","iters","dev_hist","n_halo","peers")})
except Exception as e:
    print("ERR", e)
PY
tail -n 15 gpurun_out/r02e_dist_check_n$N.err | cut -c1-400
for wl in cd27:256 powerlaw:8000000; do
for fuse in 1 0; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 5 --warmup 3 --workload $wl --no-e2e --tune dist_fuse_halo=$fuse > gpurun_out/r02e_bench_${wl/:/_}_n${N}_fuse$fuse.json 2> gpurun_out/r02e_bench_${wl/:/_}_n${N}_fuse$fuse.err; echo "bench $wl fuse=$fuse rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02e_bench_${wl/:/_}_n${N}_fuse$fuse.json") if l.startswith("{")][-1])
    print(d["config"]["workload"], "n_gpus", d["n_gpus"], "it/s %.1f"%d["value"], "ms %.2f"%d["ms_per_step"], d["config"]["iters_per_solve"], "resNorm", d["config"]["resNorm"], {k:(round(v["share"],3),v["frac_of_peak"]) for k,v in d["kernels"].items()})
except Exception as e:
    print("ERR", e)
PY
tail -n 3 gpurun_out/r02e_bench_${wl/:/_}_n${N}_fuse$fuse.err | cut -c1-300
done; done
