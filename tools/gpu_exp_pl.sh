#!/bin/bash
# power-law multi-GPU experiments: bash tools/gpu_exp_pl.sh N
N=${1:-4}
mkdir -p gpurun_out
for t in "use_pdl=1" "use_pdl=0" "dist_fuse_halo=0" "dist_peer_reduce=0"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 5 --warmup 2 --workload powerlaw:8000000 --partition nnz --no-e2e --tune $t > gpurun_out/exp_pl_n${N}_$t.json 2> gpurun_out/exp_pl_n${N}_$t.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/exp_pl_n${N}_$t.json") if l.startswith("{")][-1])
    print("$t", "it/s %.1f"%d["value"], "ms/iter %.3f"%(d["ms_per_step"]/d["config"]["iters_per_solve"]), {k:(round(v["share"],3),v["frac_of_peak"]) for k,v in d["kernels"].items()}, "sum", round(sum(v["share"] for v in d["kernels"].values()),3))
except Exception as e:
    print("$t ERR", e)
PY
done
