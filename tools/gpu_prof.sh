#!/bin/bash
# ncu evidence per /opt/skills/guides/B200_PROFILING.md: plain run first, then the launch list, then one full capture per hot kernel
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spmv_sell_kernel -s 120 -c 1 -o gpurun_out/prof_spmv -f $CMD > gpurun_out/ncu_spmv.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:vpass_kernel -s 120 -c 2 -o gpurun_out/prof_vpass -f $CMD > gpurun_out/ncu_vpass.log 2>&1
$CMD > gpurun_out/plain4.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemvn_kernel -s 170 -c 1 -o gpurun_out/prof_gemvn -f $CMD > gpurun_out/ncu_gemvn.log 2>&1
$CMD > gpurun_out/plain5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:vrow_kernel -s 130 -c 2 -o gpurun_out/prof_vrow -f $CMD > gpurun_out/ncu_vrow.log 2>&1
ls -la gpurun_out/ | tail -20
tail -n 3 gpurun_out/ncu_launches.log gpurun_out/ncu_spmv.log gpurun_out/ncu_vpass.log gpurun_out/ncu_gemvn.log gpurun_out/ncu_vrow.log | cut -c1-300
