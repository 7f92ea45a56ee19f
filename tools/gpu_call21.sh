#!/bin/bash
# round 2, call 21 (1 GPU): launch order of SELL-C-sigma slices - window order (0) / deadline-aware (1) / all longest first (2); packed tests
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_solver_gpu.py -m gpu -q -k "sigma or pack or powerlaw" --timeout 300 -p no:cacheprovider > gpurun_out/r02s_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/r02s_pytest.log | cut -c1-300
: > gpurun_out/r02s_probe_powerlaw_order.txt
for gen in powerlaw:8000000 powerlaw:1000000; do
for t in 0 1 2 0 1 2; do
  echo "== $gen sell_lpt=$t" >> gpurun_out/r02s_probe_powerlaw_order.txt
  timeout 200 python tools/spmv_probe.py --gen $gen --only packed_f32,packed_f64 --reps 15 --tune sell_lpt=$t >> gpurun_out/r02s_probe_powerlaw_order.txt 2>&1
done; done
cat gpurun_out/r02s_probe_powerlaw_order.txt | cut -c1-300
