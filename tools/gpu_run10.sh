#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_e2e.log 2>&1
tail -n 1 gpurun_out/bench_e2e.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["e2e"], d["config"]["first_solve_incl_workspace_alloc_ms"])'
timeout 900 python -m pytest tests/test_solver_gpu.py tests/test_dropin_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log | cut -c1-300
