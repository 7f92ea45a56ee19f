#!/bin/bash
# round 2, call 13 (2 GPUs): solver tests twice (MGS race fix), dist_check with the new knobs, N = 8-sized slabs on 2 ranks: late-trigger PDL on/off, one/two SpMV launches
mkdir -p gpurun_out
for i in 1 2; do timeout 900 python -m pytest tests/test_solver_gpu.py -m gpu -q > gpurun_out/r02l_pytest_$i.log 2>&1; echo "pytest $i rc=$?"; tail -n 2 gpurun_out/r02l_pytest_$i.log | cut -c1-300; done
DIST_CHECK_CASES=0,3,4,7,9,10 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py > gpurun_out/r02l_dist_check_n2.json 2> gpurun_out/r02l_dist_check_n2.err; echo "dist_check rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02l_dist_check_n2.json") if l.startswith("{")][-1])
    print("dist_check ok", d["ok"], [(c["spec"], c["orth"], c["split"], c["ok"], c["overlap_ok"]) for c in d["cases"]])
except Exception as e:
    print("ERR", e)
PY
tail -n 3 gpurun_out/r02l_dist_check_n2.err | cut -c1-300
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
for t in "use_pdl=0" "use_pdl=2" "dist_spmv_one_launch=0" "use_pdl=0"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 --workload cd27:161 --no-e2e --tune $t > gpurun_out/r02l_n2_cd27_161_$t.json 2> gpurun_out/r02l_n2_cd27_161_$t.err
show gpurun_out/r02l_n2_cd27_161_$t.json
done
for t in "use_pdl=0" "use_pdl=2"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 2 --workload powerlaw:2000000 --partition nnz --no-e2e --tune $t > gpurun_out/r02l_n2_pl2m_$t.json 2> gpurun_out/r02l_n2_pl2m_$t.err
show gpurun_out/r02l_n2_pl2m_$t.json
done
