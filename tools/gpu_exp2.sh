#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
for cfg in "cd27:256 rows use_pdl=1" "cd27:256 rows use_pdl=0" "powerlaw:8000000 nnz use_pdl=1" "powerlaw:8000000 nnz use_pdl=0"; do
set -- $cfg; wl=$1; part=$2; t=$3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 8 --warmup 3 --workload $wl --partition $part --no-e2e --tune $t > gpurun_out/exp2_n${N}_${wl/:/_}_$t.json 2> gpurun_out/exp2_n${N}_${wl/:/_}_$t.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/exp2_n${N}_${wl/:/_}_$t.json") if l.startswith("{")][-1])
    print("$wl $t", "it/s %.1f"%d["value"], "ms/iter %.4f"%(d["ms_per_step"]/d["config"]["iters_per_solve"]), {k:(round(v["share"],3),v["frac_of_peak"]) for k,v in d["kernels"].items()}, "sum", round(sum(v["share"] for v in d["kernels"].values()),3))
except Exception as e:
    print("$wl $t ERR", e)
PY
done
