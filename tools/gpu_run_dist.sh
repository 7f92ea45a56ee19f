#!/bin/bash
mkdir -p gpurun_out
NG=${1:-2}
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_solver_gpu.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29601 tools/dist_check.py > gpurun_out/dist_check.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus $NG --steps 3 --warmup 2 > gpurun_out/bench_n$NG.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_n1.log 2>&1
for f in pytest_gpu dist_check bench_n$NG bench_n1; do echo "== $f"; tail -n 5 gpurun_out/$f.log | cut -c1-2500; done
