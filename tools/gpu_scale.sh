#!/bin/bash
mkdir -p gpurun_out
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_n$n.log 2>&1
  echo "== bench_n$n"; grep "^{" gpurun_out/bench_n$n.log | tail -n 1 | cut -c1-2200; tail -n 2 gpurun_out/bench_n$n.log | grep -v "^{" | cut -c1-300
done
