#!/bin/bash
# round 2, call 2: SELL-C-sigma + packed fp64 residual
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/r02b_pytest_gpu.log
timeout 300 python tools/spmv_probe.py --gen powerlaw:8000000 > gpurun_out/r02b_probe_powerlaw.json 2>&1; echo "probe rc=$?"; cat gpurun_out/r02b_probe_powerlaw.json
timeout 300 python tools/spmv_probe.py --gen cd27:256 > gpurun_out/r02b_probe_cd27.json 2>&1; echo "probe rc=$?"; cat gpurun_out/r02b_probe_cd27.json
timeout 300 python tools/spmv_probe.py --gen lap2d:2048 > gpurun_out/r02b_probe_lap2d.json 2>&1; cat gpurun_out/r02b_probe_lap2d.json
timeout 600 python bench.py --steps 3 --warmup 2 --workload powerlaw:8000000 --no-cpu-baseline > gpurun_out/r02b_bench_powerlaw.json 2> gpurun_out/r02b_bench_powerlaw.err; echo "bench pl rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_default.json 2> gpurun_out/r02b_bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r02b_bench_powerlaw.json","gpurun_out/r02b_bench_default.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["config"]["iters_per_solve"], {k:(v["avg_ms"],v["frac_of_peak"]) for k,v in d["kernels"].items()}, d["e2e"]["value"] if d["e2e"] else None)
    except Exception as e:
        print(f, "ERR", e)
PY
tail -n 5 gpurun_out/*.err
