#!/bin/bash
# round 2, call 22 (N GPUs): power-law slabs with the deadline-aware slice order - one dist_check case, then the config [3] bench line
mkdir -p gpurun_out
N=${1:-4}
DIST_CHECK_CASES=3 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py > gpurun_out/r02t_dist_check_n$N.json 2> gpurun_out/r02t_dist_check_n$N.err; echo "dist_check rc=$?"
f=gpurun_out/r02t_bench_powerlaw_8000000_n${N}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 5 --warmup 3 --workload powerlaw:8000000 --partition nnz --no-e2e > $f.json 2> $f.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("$f.json") if l.startswith("{")][-1])
    it=d["config"]["iters_per_solve"]
    print(d["config"]["workload"], "n_gpus", d["n_gpus"], "it/s %.1f"%d["value"], "ms %.2f"%d["ms_per_step"], it, "resNorm", d["config"]["resNorm"])
    print("   ", {k:(round(1e3*v["ms_total"]/d["steps"]/it,1),v["frac_of_peak"]) for k,v in d["kernels"].items()})
except Exception as e:
    print("ERR", e)
PY
tail -n 3 $f.err | grep -v "OMP_NUM_THREADS\|^\*\*\*" | cut -c1-300
