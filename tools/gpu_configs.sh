#!/bin/bash
# BASELINE.json configs[0..4] on one B200 (configs[2] is the default bench)
mkdir -p gpurun_out
timeout 600 python bench.py --workload lap2d:512 --mode baseline --rlen 50 --steps 5 --warmup 3 --cpu-sample lap2d:512 > gpurun_out/bench_c1.log 2>&1
timeout 600 python bench.py --workload lap2d:2048 --mode mixed --rlen 50 --steps 5 --warmup 3 --cpu-sample lap2d:1024 > gpurun_out/bench_c2.log 2>&1
timeout 900 python bench.py --workload powerlaw:8000000 --mode mixed --rlen 100 --steps 3 --warmup 2 --cpu-sample powerlaw:500000 > gpurun_out/bench_c4.log 2>&1
timeout 900 python tools/kernel_perf_test.py --gen cd27:256 --cpu --cpu-gen cd27:64 > gpurun_out/kernel_perf_c5.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c3.log 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
for f in bench_c1 bench_c2 bench_c4 kernel_perf_c5 bench_c3 bench_ref; do echo "== $f"; tail -n 2 gpurun_out/$f.log | cut -c1-1500; done
