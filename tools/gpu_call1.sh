#!/bin/bash
# round 2, call 1: GPU tests, default bench (with the real-size CPU baseline + parity block), power-law evidence
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r02_pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 3 --warmup 2 --workload powerlaw:8000000 --no-cpu-baseline > gpurun_out/r02_bench_powerlaw_before.json 2> gpurun_out/r02_bench_powerlaw_before.err; echo "bench pl rc=$?"
timeout 300 python tools/spmv_probe.py --gen powerlaw:8000000 > gpurun_out/r02_probe_powerlaw_before.json 2>&1; echo "probe rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_tile_kernel -c 2 -o gpurun_out/r02_prof_spmv_tile_powerlaw -f python tools/spmv_probe.py --gen powerlaw:8000000 --reps 1 --only csr_f32 > gpurun_out/r02_ncu_spmv_tile_powerlaw.log 2>&1; echo "ncu rc=$?"
cat gpurun_out/r02_probe_powerlaw_before.json
head -c 3000 gpurun_out/r02_bench_default.json
