#!/bin/bash
mkdir -p gpurun_out
NG=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29601 tools/dist_check.py > gpurun_out/dist_check_n$NG.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus $NG --steps 5 --warmup 3 > gpurun_out/bench_n$NG.log 2>&1
for f in dist_check_n$NG bench_n$NG; do echo "== $f"; grep "^{" gpurun_out/$f.log | tail -n 1 | cut -c1-2500; tail -n 2 gpurun_out/$f.log | grep -v "^{" | cut -c1-300; done
