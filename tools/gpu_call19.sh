#!/bin/bash
# round 2, call 19 (N GPUs, default 4): partitioned path of the committed tree - dist_check (plan bits, halo, knob invariance), then the
# bench line of cd27:256 (row split, with e2e) and powerlaw:8000000 (nnz split)
mkdir -p gpurun_out
N=${1:-4}
DIST_CHECK_CASES=${CASES:-0,3,7,10} timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py > gpurun_out/r02q_dist_check_n$N.json 2> gpurun_out/r02q_dist_check_n$N.err; echo "dist_check rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02q_dist_check_n$N.json") if l.startswith("{")][-1])
    print("dist_check ok", d["ok"], [(c["spec"], c["orth"], c["split"], c["ok"], c["overlap_ok"], c["peers"], c["n_halo"]) for c in d["cases"]])
except Exception as e:
    print("ERR", e)
PY
tail -n 4 gpurun_out/r02q_dist_check_n$N.err | cut -c1-300
for cfg in "cd27:256 rows --e2e-steps=1" "powerlaw:8000000 nnz --no-e2e"; do
set -- $cfg; wl=$1; part=$2; extra=$3
f=gpurun_out/r02q_bench_${wl/:/_}_n${N}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 5 --warmup 3 --workload $wl --partition $part $extra > $f.json 2> $f.err; echo "bench $wl rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("$f.json") if l.startswith("{")][-1])
    print(d["config"]["workload"], "n_gpus", d["n_gpus"], "it/s %.1f"%d["value"], "ms %.2f"%d["ms_per_step"], d["config"]["iters_per_solve"], "resNorm", d["config"]["resNorm"], "e2e", d["e2e"]["value"] if d["e2e"] else None)
    print("   ", {k:(round(v["share"],3),v["frac_of_peak"],v["launches"]) for k,v in d["kernels"].items()})
except Exception as e:
    print("ERR", e)
PY
tail -n 3 $f.err | grep -v "OMP_NUM_THREADS\|^\*\*\*" | cut -c1-300
done
