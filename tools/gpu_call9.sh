#!/bin/bash
# round 2, call 9 (2 GPUs): N = 8-sized slabs (2.1 M rows per rank) on 2 ranks - PDL on/off, fused halo on/off - against the same slab on ONE GPU;
# correctness of the re-ordered tail / single-fence push; host cast probe
mkdir -p gpurun_out
gcc -O3 -march=native -fopenmp tools/probes/hostcast.c -o /tmp/hostcast && /tmp/hostcast > gpurun_out/r02h_hostcast.txt 2>&1; tail -n 8 gpurun_out/r02h_hostcast.txt
DIST_CHECK_CASES=0,3,4,7,9,10 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py > gpurun_out/r02h_dist_check_n2.json 2> gpurun_out/r02h_dist_check_n2.err; echo "dist_check rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02h_dist_check_n2.json") if l.startswith("{")][-1])
    print("dist_check ok", d["ok"], [(c["spec"], c["orth"], c["split"], c["ok"]) for c in d["cases"]])
except Exception as e:
    print("ERR", e)
PY
show() {
python - <<PY
import json
try:
    d=json.loads([l for l in open("$1") if l.startswith("{")][-1])
    it=d["config"]["iters_per_solve"]
    print("$1".split("/")[-1], "it/s %.1f"%d["value"], "us/iter %.1f"%(1e3*d["ms_per_step"]/it), {k:(round(1e3*v["ms_total"]/d["steps"]/it,1),v["frac_of_peak"]) for k,v in d["kernels"].items()})
    if d.get("kernels_ms_per_rank"): print("    per rank:", d["kernels_ms_per_rank"])
except Exception as e:
    print("$1 ERR", e)
PY
}
for t in "use_pdl=1" "use_pdl=0" "dist_fuse_halo=0"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 --workload cd27:161 --no-e2e --tune $t > gpurun_out/r02h_n2_cd27_161_$t.json 2> gpurun_out/r02h_n2_cd27_161_$t.err
show gpurun_out/r02h_n2_cd27_161_$t.json
done
for t in "use_pdl=1" "use_pdl=0"; do
timeout 300 python bench.py --steps 10 --warmup 3 --workload cd27:128 --no-e2e --no-cpu-baseline --no-multi-restart --tune $t > gpurun_out/r02h_n1_cd27_128_$t.json 2> gpurun_out/r02h_n1_cd27_128_$t.err
show gpurun_out/r02h_n1_cd27_128_$t.json
done
for t in "use_pdl=1" "use_pdl=0"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 2 --workload powerlaw:2000000 --partition nnz --no-e2e --tune $t > gpurun_out/r02h_n2_pl2m_$t.json 2> gpurun_out/r02h_n2_pl2m_$t.err
show gpurun_out/r02h_n2_pl2m_$t.json
done
