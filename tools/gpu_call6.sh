#!/bin/bash
# round 2, call 6 (1 GPU): full GPU suite, kernel sweep (configs[4]), small configs
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02f_pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -n 12 gpurun_out/r02f_pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --workload sweep --steps 5 > gpurun_out/r02f_bench_sweep.json 2> gpurun_out/r02f_bench_sweep.err; echo "sweep rc=$?"; tail -n 3 gpurun_out/r02f_bench_sweep.err
timeout 600 python bench.py --impl reference --workload sweep > gpurun_out/r02f_bench_sweep_ref.json 2> gpurun_out/r02f_bench_sweep_ref.err; echo "sweep ref rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r02f_bench_sweep.json").read().strip().splitlines()[-1])
    print("value", d["value"], "cpu", d["cpu_baseline"]["value"] if d["cpu_baseline"] else None, d["config"]["wall_s"])
    for key, t in d["sweep"].items():
        print(key, t["spmv"])
        for k, v in t["add_vector"].items():
            print("   ", k, v)
    c=d["cpu_baseline"]["sweep"]
    print("cpu", c["spmv"])
    for k, v in c["add_vector"].items():
        print("   ", k, v)
except Exception as e:
    print("ERR", e)
PY
for wl in lap2d:512 lap2d:2048; do
mode=mixed; [ $wl = lap2d:512 ] && mode=baseline
timeout 300 python bench.py --steps 10 --warmup 3 --workload $wl --rlen 50 --mode $mode --no-cpu-baseline > gpurun_out/r02f_bench_${wl/:/_}.json 2> gpurun_out/r02f_bench_${wl/:/_}.err; echo "bench $wl rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02f_bench_${wl/:/_}.json").read().strip().splitlines()[-1])
    print(d["config"]["workload"], d["config"]["mode"], "it/s %.1f"%d["value"], "ms %.3f"%d["ms_per_step"], d["config"]["iters_per_solve"], "launches", d["gpu_launches"], {k:(v["avg_ms"],v["frac_of_peak"]) for k,v in d["kernels"].items()}, "e2e", d["e2e"]["value"])
except Exception as e:
    print("ERR", e)
PY
done
