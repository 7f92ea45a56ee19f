#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider -k "add_vector or gemv or fused" > gpurun_out/pytest_ops.log 2>&1
tail -n 3 gpurun_out/pytest_ops.log
timeout 600 python tools/tune.py --no-spmv --ks 1,4,8,9,12,16,33,40,48,49,56,64,65 --variants default,direct16,vrow48,vrow64 > gpurun_out/tune_vrow64_16m.txt 2>&1
cat gpurun_out/tune_vrow64_16m.txt
timeout 600 python tools/tune.py --no-spmv --n 2097152 --ks 1,4,8,9,12,16,33,40,48,49,56,64,65 --variants default,direct16,vrow48,vrow64 > gpurun_out/tune_vrow64_2m.txt 2>&1
cat gpurun_out/tune_vrow64_2m.txt
