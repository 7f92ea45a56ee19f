#!/bin/bash
# round 2, call 14 (1 GPU): full GPU suite, smoke, default bench (all legs)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02m_pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -n 5 gpurun_out/r02m_pytest_gpu.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as e; e.smoke(); print('smoke ok')" > gpurun_out/r02m_smoke.log 2>&1; tail -n 2 gpurun_out/r02m_smoke.log
timeout 900 python bench.py > gpurun_out/r02m_bench_default.json 2> gpurun_out/r02m_bench_default.err; echo "bench rc=$?"; tail -n 1 gpurun_out/r02m_bench_default.json | cut -c1-900
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02m_bench_reference.json 2> gpurun_out/r02m_bench_reference.err; echo "ref rc=$?"; tail -n 1 gpurun_out/r02m_bench_reference.json | cut -c1-600
