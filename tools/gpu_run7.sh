#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
tail -n 6 gpurun_out/pytest_gpu.log | cut -c1-600
for w in cd27:128 lap2d:512; do
  extra=""; [ $w = lap2d:512 ] && extra="--mode baseline --rlen 50"
  for pdl in 1 0; do
    timeout 900 python bench.py --tune use_pdl=$pdl --workload $w $extra --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_${w/:/_}_pdl$pdl.log 2>&1
    echo "== $w pdl=$pdl"; tail -n 1 gpurun_out/bench_${w/:/_}_pdl$pdl.log | cut -c1-200
  done
done
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_n1.log 2>&1
tail -n 1 gpurun_out/bench_n1.log | cut -c1-300
