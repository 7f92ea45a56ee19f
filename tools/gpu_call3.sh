#!/bin/bash
# round 2, call 3: packed-kernel variants on power-law / stencil matrices, substrate test, ncu of the sigma kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_substrate_gpu.py tests/test_ops_gpu.py -m gpu -x -q > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/r02c_pytest.log
: > gpurun_out/r02c_tune_sell.txt
for gen in powerlaw:8000000 cd27:256; do
for v in 0 1 2 3; do for b in 128 256; do
  echo "== $gen sell_variant=$v sell_block=$b" >> gpurun_out/r02c_tune_sell.txt
  timeout 200 python tools/spmv_probe.py --gen $gen --only packed_f32,packed_f64 --reps 7 --tune sell_variant=$v --tune sell_block=$b >> gpurun_out/r02c_tune_sell.txt 2>&1
done; done; done
cat gpurun_out/r02c_tune_sell.txt | grep -v "^$" | cut -c1-400
timeout 200 python tools/spmv_probe.py --gen powerlaw:8000000 --reps 1 --only packed_f32 > gpurun_out/plain_probe.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_sell_kernel -c 2 -o gpurun_out/r02c_prof_spmv_sell_powerlaw -f python tools/spmv_probe.py --gen powerlaw:8000000 --reps 1 --only packed_f32 > gpurun_out/r02c_ncu_spmv_sell_powerlaw.log 2>&1; echo "ncu rc=$?"
