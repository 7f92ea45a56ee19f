#!/bin/bash
mkdir -p gpurun_out
for pdl in 1 0; do
timeout 600 python bench.py --tune use_pdl=$pdl --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_e2e_pdl$pdl.log 2>&1
echo "pdl=$pdl"; tail -n 1 gpurun_out/bench_e2e_pdl$pdl.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["e2e"], d["config"]["first_solve_incl_workspace_alloc_ms"])'
done
