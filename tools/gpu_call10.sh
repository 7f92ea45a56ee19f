#!/bin/bash
# round 2, call 10 (1 GPU): new host-path tests, e2e overlapped vs serial, PDL threshold on one GPU, host cast probe with all threads
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_solver_gpu.py tests/test_ops_gpu.py -m gpu -q -x > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/r02i_pytest.log | cut -c1-400
gcc -O3 -march=native -fopenmp tools/probes/hostcast.c -o /tmp/hostcast && OMP_NUM_THREADS=$(nproc) /tmp/hostcast > gpurun_out/r02i_hostcast.txt 2>&1; tail -n 8 gpurun_out/r02i_hostcast.txt
show() {
python - <<PY
import json
try:
    d=json.loads([l for l in open("$1") if l.startswith("{")][-1])
    it=d["config"]["iters_per_solve"]
    print("$1".split("/")[-1], "it/s %.1f"%d["value"], "us/iter %.1f"%(1e3*d["ms_per_step"]/it), "e2e", d["e2e"])
except Exception as e:
    print("$1 ERR", e)
PY
}
for t in "host_overlap=1" "host_overlap=0" "host_threads=8"; do
MPG_TRACE=0 timeout 600 python bench.py --steps 3 --warmup 3 --e2e-steps 3 --no-cpu-baseline --no-multi-restart --tune $t > gpurun_out/r02i_e2e_$t.json 2> gpurun_out/r02i_e2e_$t.err
show gpurun_out/r02i_e2e_$t.json
done
for wl in "lap2d:512 baseline 50" "cd27:64 mixed 100" "cd27:100 mixed 100" "lap2d:1024 mixed 50"; do
set -- $wl
for t in "use_pdl=2" "use_pdl=0"; do
timeout 300 python bench.py --steps 10 --warmup 3 --workload $1 --mode $2 --rlen $3 --no-e2e --no-cpu-baseline --no-multi-restart --tune $t > gpurun_out/r02i_pdl_${1/:/_}_$t.json 2> gpurun_out/r02i_pdl_${1/:/_}_$t.err
show gpurun_out/r02i_pdl_${1/:/_}_$t.json
done
done
