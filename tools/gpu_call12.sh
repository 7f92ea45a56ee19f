#!/bin/bash
# round 2, call 12 (1 GPU): late PDL trigger (use_pdl=2: attribute on every launch) vs no PDL; solver tests on the new triggers
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_solver_gpu.py -m gpu -q -x > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r02k_pytest.log | cut -c1-400
show() {
python - <<PY
import json
try:
    d=json.loads([l for l in open("$1") if l.startswith("{")][-1])
    it=d["config"]["iters_per_solve"]
    print("$1".split("/")[-1], "it/s %.1f"%d["value"], "us/iter %.1f"%(1e3*d["ms_per_step"]/it), d["config"]["step_ms_min_med_max"])
except Exception as e:
    print("$1 ERR", e)
PY
}
for wl in "lap2d:512 baseline 50" "cd27:64 mixed 100" "cd27:100 mixed 100" "cd27:128 mixed 100" "lap2d:2048 mixed 50" "cd27:256 mixed 100"; do
set -- $wl
for t in "use_pdl=2" "use_pdl=0" "use_pdl=2"; do
timeout 300 python bench.py --steps 10 --warmup 3 --workload $1 --mode $2 --rlen $3 --no-e2e --no-cpu-baseline --no-multi-restart --tune $t > gpurun_out/r02k_pdl_${1/:/_}_$t.json 2> gpurun_out/r02k_pdl_${1/:/_}_$t.err
show gpurun_out/r02k_pdl_${1/:/_}_$t.json
done
done
