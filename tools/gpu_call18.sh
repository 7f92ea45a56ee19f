#!/bin/bash
# round 2, call 18 (1 GPU): counter evidence for the power-law SpMV - ncu --set full of the CSR tile kernel and of the shipped packed
# (SELL-C-sigma) kernel on powerlaw:8000000, each after the same command exited 0 without ncu; launch list of the default bench
mkdir -p gpurun_out
P="python tools/spmv_probe.py --gen powerlaw:8000000 --reps 1"
timeout 200 $P --only csr_f32,packed_f32 > gpurun_out/r02p_probe_powerlaw.json 2> gpurun_out/r02p_probe_powerlaw.err && {
cat gpurun_out/r02p_probe_powerlaw.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:spmv_tile_kernel -c 2 -f -o gpurun_out/r02p_prof_spmv_tile_powerlaw $P --only csr_f32 > gpurun_out/r02p_ncu_tile.log 2>&1; echo "ncu tile rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:spmv_sell_kernel -c 2 -f -o gpurun_out/r02p_prof_spmv_sell_powerlaw $P --only packed_f32 > gpurun_out/r02p_ncu_sell.log 2>&1; echo "ncu sell rc=$?"
}
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-multi-restart"
timeout 300 $CMD > gpurun_out/r02p_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02p_ncu_launches_cd27_256.csv $CMD > gpurun_out/r02p_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
wc -l gpurun_out/r02p_ncu_launches_cd27_256.csv
ls -la gpurun_out/*.ncu-rep
