#!/bin/bash
# round 2, call 16 (2 GPUs): flag-in-data all-reduce vs fence + flag: correctness (dist_check incl. knob invariance), primitive latencies, N = 8-sized slabs
mkdir -p gpurun_out
DIST_CHECK_CASES=0,1,3,4,5,7,9,10 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py > gpurun_out/r02o_dist_check_n2.json 2> gpurun_out/r02o_dist_check_n2.err; echo "dist_check rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02o_dist_check_n2.json") if l.startswith("{")][-1])
    print("dist_check ok", d["ok"], [(c["spec"], c["orth"], c["split"], c["ok"], c["overlap_ok"]) for c in d["cases"]])
except Exception as e:
    print("ERR", e)
PY
tail -n 3 gpurun_out/r02o_dist_check_n2.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 tools/dist_latency.py cd27:128 > gpurun_out/r02o_dist_latency_n2.json 2> gpurun_out/r02o_dist_latency_n2.err; grep "^{" gpurun_out/r02o_dist_latency_n2.json | tail -n 1
show() {
python - <<PY
import json
try:
    d=json.loads([l for l in open("$1") if l.startswith("{")][-1])
    it=d["config"]["iters_per_solve"]
    print("$1".split("/")[-1], "it/s %.1f"%d["value"], "us/iter %.1f"%(1e3*d["ms_per_step"]/it), {k:(round(1e3*v["ms_total"]/d["steps"]/it,1),v["frac_of_peak"]) for k,v in d["kernels"].items()})
except Exception as e:
    print("$1 ERR", e)
PY
}
for t in "dist_ll_reduce=1" "dist_ll_reduce=0" "dist_ll_reduce=1" "dist_ll_reduce=0"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 --workload cd27:161 --no-e2e --tune $t > gpurun_out/r02o_n2_cd27_161_$t.json 2> gpurun_out/r02o_n2_cd27_161_$t.err
show gpurun_out/r02o_n2_cd27_161_$t.json
done
