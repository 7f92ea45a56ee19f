"""bench.py contract checks that need no GPU: the reference arm (the reference's own CPU path from oracle/_ref) prints one
well-formed JSON line, under torchrun only rank 0 prints, and the product arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "gmres_perf_test")
ARGS = ["--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "lap2d:48", "--rlen", "20"]


def _json_lines(out):
    return [json.loads(l) for l in out.splitlines() if l.startswith("{")]


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built (needs /root/reference at build time)")
def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + ARGS, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = _json_lines(out.stdout)
    assert len(lines) == 1
    d = lines[0]
    assert d["impl"] == "reference" and d["metric"] == "gmres_ir_inner_iterations_per_sec" and d["unit"] == "it/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"] == "lap2d:48"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and "lap2d:48 at full size" in cb["sample"]
    assert d["config"]["n_rows"] == 48 * 48 and d["config"]["iters_per_solve"] > 0 and d["config"]["resNorm"] > 0
    # ms_per_step x steps is what was actually timed (no extrapolation)
    assert abs(d["ms_per_step"] * d["steps"] * 1e-3 * d["value"] - d["config"]["iters_per_solve"] * d["steps"]) < 1e-6 * d["config"]["iters_per_solve"] + 1e-9
    assert d["e2e"] == {"value": d["value"], "unit": "it/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built")
def test_reference_arm_under_torchrun_only_rank0_prints():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29713",
           os.path.join(ROOT, "bench.py"), "--gpus", "2"] + ARGS
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = _json_lines(out.stdout)
    assert len(lines) == 1 and lines[0]["impl"] == "reference" and lines[0]["n_gpus"] == 2
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm must not inherit that
    ncpu = len(os.sched_getaffinity(0))
    assert lines[0]["cpu_baseline"]["cores"] == ncpu or ncpu == 1


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built")
def test_reference_arm_caps_steps_and_reports_what_it_ran():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "20", "--warmup", "5", "--workload", "lap2d:48",
                          "--rlen", "20"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert out.returncode == 0, out.stderr[-2000:]
    d = _json_lines(out.stdout)[0]
    assert d["steps"] == 2 and d["warmup"] == 0 and d["steps_requested"] == 20 and d["warmup_requested"] == 5
    ncpu = len(os.sched_getaffinity(0))
    assert d["cpu_baseline"]["cores"] == ncpu or ncpu == 1


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--workload", "lap2d:16"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0
    assert "no CUDA device" in (out.stdout + out.stderr)
    assert not _json_lines(out.stdout)
