#!/bin/bash
# full single-GPU validation: all GPU tests, then the default bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -n 8 gpurun_out/pytest_gpu.log | cut -c1-600
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.log 2>&1
tail -n 1 gpurun_out/bench_n1.log | cut -c1-4000
