#!/bin/bash
# round 2, call 11 (1 GPU): host-path tests after the pool_free fix, e2e overlapped vs serial, tail kernel at slab size
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_solver_gpu.py tests/test_ops_gpu.py -m gpu -q > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/r02j_pytest.log | cut -c1-400
show() {
python - <<PY
import json
try:
    d=json.loads([l for l in open("$1") if l.startswith("{")][-1])
    it=d["config"]["iters_per_solve"]
    print("$1".split("/")[-1], "it/s %.1f"%d["value"], "us/iter %.1f"%(1e3*d["ms_per_step"]/it), {k:round(1e3*v["ms_total"]/d["steps"]/it,1) for k,v in d["kernels"].items()}, "e2e", d["e2e"])
except Exception as e:
    print("$1 ERR", e)
PY
}
for t in "host_overlap=1" "host_overlap=0"; do
timeout 600 python bench.py --steps 3 --warmup 3 --e2e-steps 3 --no-cpu-baseline --no-multi-restart --tune $t > gpurun_out/r02j_e2e_$t.json 2> gpurun_out/r02j_e2e_$t.err
show gpurun_out/r02j_e2e_$t.json
done
timeout 300 python bench.py --steps 10 --warmup 3 --workload cd27:128 --no-e2e --no-cpu-baseline --no-multi-restart > gpurun_out/r02j_n1_cd27_128.json 2> gpurun_out/r02j_n1_cd27_128.err
show gpurun_out/r02j_n1_cd27_128.json
timeout 300 python bench.py --steps 10 --warmup 3 --workload lap2d:2048 --rlen 50 > gpurun_out/r02j_lap2d_2048.json 2> gpurun_out/r02j_lap2d_2048.err
show gpurun_out/r02j_lap2d_2048.json
