#!/bin/bash
# round 2, call 4: ILU(0)-Jacobi, loaders, full GPU suite, benches
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ilu_gpu.py -m gpu -x -q > gpurun_out/r02d_pytest_ilu.log 2>&1; echo "pytest ilu rc=$?"; tail -n 25 gpurun_out/r02d_pytest_ilu.log | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02d_pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -n 15 gpurun_out/r02d_pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 2 --workload powerlaw:8000000 --no-cpu-baseline > gpurun_out/r02d_bench_powerlaw.json 2> gpurun_out/r02d_bench_powerlaw.err; echo "bench pl rc=$?"
timeout 300 python gmres_perf_test.py --gen cd27:128 --rlen 50 --orth cgsr --prec ilu_jacobi --jacobi-steps 3 --tol 1e-9 --json > gpurun_out/r02d_ilu_cd27_128.log 2>&1; tail -n 4 gpurun_out/r02d_ilu_cd27_128.log
timeout 300 python gmres_perf_test.py --gen cd27:128 --rlen 50 --orth cgsr --prec identity --tol 1e-9 --json > gpurun_out/r02d_id_cd27_128.log 2>&1; tail -n 4 gpurun_out/r02d_id_cd27_128.log
python - <<'PY'
import json
for f in ("gpurun_out/r02d_bench_powerlaw.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["config"]["iters_per_solve"], {k:(v["avg_ms"],v["frac_of_peak"]) for k,v in d["kernels"].items()}, d["e2e"]["value"] if d["e2e"] else None)
    except Exception as e:
        print(f, "ERR", e)
PY
