#!/bin/bash
# round 2, call 8 (8 GPUs): correctness + both partitioned workloads + primitive latencies at N = 8
export EXTRA="--no-e2e"
bash tools/gpu_scale.sh 8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29613 tools/dist_latency.py cd27:256 > gpurun_out/r02_dist_latency_n8.json 2> gpurun_out/r02_dist_latency_n8.err; echo "latency rc=$?"
grep "^{" gpurun_out/r02_dist_latency_n8.json | tail -n 1
