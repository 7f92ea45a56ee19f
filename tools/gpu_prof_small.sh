#!/bin/bash
# launch list (per-kernel GPU durations) of the slab-sized problem: separates kernel time from launch gaps
mkdir -p gpurun_out
CMD="python bench.py --workload cd27:128 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_small.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_small.csv $CMD > gpurun_out/ncu_launches_small.log 2>&1
tail -n 2 gpurun_out/plain_small.log | cut -c1-600; tail -n 3 gpurun_out/ncu_launches_small.log | cut -c1-300; wc -l gpurun_out/launches_small.csv
