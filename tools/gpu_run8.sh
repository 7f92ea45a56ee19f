#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
tail -n 6 gpurun_out/pytest_gpu.log | cut -c1-600
for ft in 1 0; do
timeout 600 python bench.py --tune fuse_tail=$ft --workload lap2d:512 --mode baseline --rlen 50 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c1_ft$ft.log 2>&1
echo "c1 fuse_tail=$ft"; tail -n 1 gpurun_out/bench_c1_ft$ft.log | cut -c1-200
timeout 600 python bench.py --tune fuse_tail=$ft --workload cd27:128 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_128_ft$ft.log 2>&1
echo "cd27:128 fuse_tail=$ft"; tail -n 1 gpurun_out/bench_128_ft$ft.log | cut -c1-200
done
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1
tail -n 1 gpurun_out/bench_n1.log | cut -c1-300
