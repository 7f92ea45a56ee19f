#!/bin/bash
# round 2, call 20 (1 GPU): longest-first slice order of SELL-C-sigma plans - full GPU test suite + smoke on the new library, A/B of the
# packed power-law SpMV (sell_lpt = 1 / 0), config [3] bench line
mkdir -p gpurun_out
time timeout 600 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/r02r_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -n 6 gpurun_out/r02r_pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/r02r_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r02r_smoke.log | cut -c1-300
: > gpurun_out/r02r_probe_powerlaw_lpt.txt
for t in 1 0 1 0; do
  echo "== sell_lpt=$t" >> gpurun_out/r02r_probe_powerlaw_lpt.txt
  timeout 200 python tools/spmv_probe.py --gen powerlaw:8000000 --only packed_f32,packed_f64 --reps 9 --tune sell_lpt=$t >> gpurun_out/r02r_probe_powerlaw_lpt.txt 2>&1
done
cat gpurun_out/r02r_probe_powerlaw_lpt.txt | cut -c1-300
timeout 400 python bench.py --workload powerlaw:8000000 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r02r_bench_powerlaw.json 2> gpurun_out/r02r_bench_powerlaw.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02r_bench_powerlaw.json") if l.startswith("{")][-1])
print("it/s %.1f"%d["value"], "ms %.2f"%d["ms_per_step"], d["config"]["iters_per_solve"], {k:(v["avg_ms"],v["frac_of_peak"]) for k,v in d["kernels"].items()}, "e2e", d["e2e"] and d["e2e"]["value"])
PY
