#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_solver_gpu.py tests/test_dropin_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -n 30 gpurun_out/pytest_gpu.log | cut -c1-600
