#!/bin/bash
mkdir -p gpurun_out
NG=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29601 tools/dist_check.py > gpurun_out/dist_check_n$NG.log 2>&1
for n in 8 4; do
  if [ $n -le $NG ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_n$n.log 2>&1
  fi
done
for f in dist_check_n$NG bench_n8 bench_n4; do echo "== $f"; grep "^{" gpurun_out/$f.log | cut -c1-1800; tail -n 2 gpurun_out/$f.log | grep -v "^{" | cut -c1-400; done
