#!/bin/bash
# round 2 ncu evidence (B200_PROFILING.md): plain run first, then the launch list, then full captures of the hot kernels
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-multi-restart"
$CMD > gpurun_out/r02_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_ncu_launches_cd27_256.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/r02_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:spmv_sell_kernel -s 101 -c 3 -o gpurun_out/r02_prof_spmv_sell_cd27 -f $CMD > gpurun_out/r02_ncu_spmv_sell.log 2>&1
echo "spmv capture rc=$?"
$CMD > gpurun_out/r02_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:vpass_kernel -s 120 -c 2 -o gpurun_out/r02_prof_vpass -f $CMD > gpurun_out/r02_ncu_vpass.log 2>&1
echo "vpass capture rc=$?"
ls -la gpurun_out/ | grep r02_ | tail -12
