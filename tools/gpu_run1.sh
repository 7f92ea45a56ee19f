#!/bin/bash
# first GPU session: parity tests, smoke (+memcheck), small and full bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
python -c "import __graft_entry__ as e; e.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_ops.log 2>&1
timeout 900 python -m pytest tests/test_solver_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_solver.log 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py --workload cd27:64 --steps 2 --warmup 1 --cpu-sample cd27:32 > gpurun_out/bench_small.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 2 --cpu-sample cd27:96 > gpurun_out/bench_full.log 2>&1
for f in pytest_ops pytest_solver smoke bench_small bench_full; do echo "== $f"; tail -n 4 gpurun_out/$f.log; done
