#!/bin/bash
# round 2, call 7 (1 GPU): full GPU suite, smoke, default bench, ncu launch list + full captures (tools/gpu_prof_r02.sh)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02g_pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -n 12 gpurun_out/r02g_pytest_gpu.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as e; e.smoke(); print('smoke ok')" > gpurun_out/r02g_smoke.log 2>&1; tail -n 2 gpurun_out/r02g_smoke.log
timeout 900 python bench.py > gpurun_out/r02g_bench_default.json 2> gpurun_out/r02g_bench_default.err; echo "bench rc=$?"; tail -n 1 gpurun_out/r02g_bench_default.json | cut -c1-1500
bash tools/gpu_prof_r02.sh
