#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_solver_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
tail -n 12 gpurun_out/pytest_gpu.log | cut -c1-400
for spec in cd27:256 lap2d:2048 cd27:128; do
  timeout 600 python tools/tune.py --ks 4 --variants default --spec $spec --n 1048576 > gpurun_out/tune_sell_${spec/:/_}.txt 2>&1
  echo "== $spec"; grep -v "^ \|^---\|^n=\|^sum\|k1" gpurun_out/tune_sell_${spec/:/_}.txt | tail -8
done
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1
tail -n 1 gpurun_out/bench_n1.log | cut -c1-3000
