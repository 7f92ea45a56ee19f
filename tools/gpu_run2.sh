#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_ops.log 2>&1
timeout 1500 python tools/tune.py --no-spmv > gpurun_out/tune.log 2>&1
tail -n 3 gpurun_out/pytest_ops.log; cat gpurun_out/tune.log
