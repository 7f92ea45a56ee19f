#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_ops.log 2>&1
timeout 900 python -m pytest tests/test_solver_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_solver.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --cpu-sample cd27:96 > gpurun_out/bench_full.log 2>&1
for f in pytest_ops pytest_solver smoke bench_full; do echo "== $f"; tail -n 6 gpurun_out/$f.log; done
