#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29604 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench_n4.log 2>&1
grep "^{" gpurun_out/bench_n4.log | cut -c1-300
