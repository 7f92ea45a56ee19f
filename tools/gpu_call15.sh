#!/bin/bash
# round 2, call 15 (2 GPUs): halo push in the head of the SpMV kernel vs in the Arnoldi tail, on N = 8-sized slabs; correctness first
mkdir -p gpurun_out
DIST_CHECK_CASES=0,2,3,4,7,8,9,10 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py > gpurun_out/r02n_dist_check_n2.json 2> gpurun_out/r02n_dist_check_n2.err; echo "dist_check rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02n_dist_check_n2.json") if l.startswith("{")][-1])
    print("dist_check ok", d["ok"], [(c["spec"], c["orth"], c["split"], c["ok"], c["overlap_ok"]) for c in d["cases"]])
except Exception as e:
    print("ERR", e)
PY
tail -n 3 gpurun_out/r02n_dist_check_n2.err | cut -c1-300
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
for t in "dist_push_in_spmv=1" "dist_push_in_spmv=0" "dist_push_in_spmv=1"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 --workload cd27:161 --no-e2e --tune $t > gpurun_out/r02n_n2_cd27_161_$t.json 2> gpurun_out/r02n_n2_cd27_161_$t.err
show gpurun_out/r02n_n2_cd27_161_$t.json
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 --workload cd27:256 > gpurun_out/r02n_n2_cd27_256.json 2> gpurun_out/r02n_n2_cd27_256.err
show gpurun_out/r02n_n2_cd27_256.json
