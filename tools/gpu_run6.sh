#!/bin/bash
mkdir -p gpurun_out
python tools/vpass_timeline.py > gpurun_out/vpass_timeline.txt 2>&1; cat gpurun_out/vpass_timeline.txt
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -n 6 gpurun_out/pytest_gpu.log | cut -c1-600
timeout 900 python bench.py --workload cd27:128 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_cd27_128.log 2>&1
tail -n 1 gpurun_out/bench_cd27_128.log | cut -c1-400
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1
tail -n 1 gpurun_out/bench_n1.log | cut -c1-400
