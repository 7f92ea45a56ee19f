#!/bin/bash
# round 2, call 17 (1 GPU): full validation of the tree as committed - smoke(), default bench (with the full-size CPU baseline +
# parity block), all GPU tests, the reference arm
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as e; e.smoke()" > gpurun_out/r02m_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r02m_smoke.log | cut -c1-300
time timeout 600 python bench.py > gpurun_out/r02m_bench_default.json 2> gpurun_out/r02m_bench_default.err; echo "bench rc=$?"
tail -n 1 gpurun_out/r02m_bench_default.json | cut -c1-6000
time timeout 720 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider --durations=15 > gpurun_out/r02m_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -n 25 gpurun_out/r02m_pytest_gpu.log | cut -c1-300
time timeout 300 python bench.py --impl reference > gpurun_out/r02m_bench_reference.json 2> gpurun_out/r02m_bench_reference.err; echo "ref rc=$?"
tail -n 1 gpurun_out/r02m_bench_reference.json | cut -c1-3000
