"""Markdown tables for DESIGN.md from the bench lines committed under profiles/.

    python tools/make_tables.py scaling profiles/r02_scale_cd27_256_n{1,2,4,8}.json
    python tools/make_tables.py sweep profiles/r02_bench_sweep.json
"""
import json
import sys


def last_line(path):
    return json.loads([l for l in open(path) if l.startswith("{")][-1])


def scaling(paths):
    rows = [last_line(p) for p in paths]
    rows.sort(key=lambda d: d["n_gpus"])
    base = rows[0]["value"] / rows[0]["n_gpus"]
    print("| GPUs | it/s | time to solution | µs / iteration | efficiency vs N = 1 | SpMV | V passes | gemv-N | tail + push (`elementwise`) |")
    print("|---|---|---|---|---|---|---|---|---|")
    for d in rows:
        it = d["config"]["iters_per_solve"]
        k = d["kernels"]
        f = lambda c: f"{k[c]['frac_of_peak']:.2f}" if c in k else "—"
        print(f"| {d['n_gpus']} | {d['value']:.1f} | {d['ms_per_step']:.1f} ms | {1e3 * d['ms_per_step'] / it:.1f} | {d['value'] / (base * d['n_gpus']):.3f} | "
              f"{f('spmv_f32')} | {f('vpass')} | {f('gemvn')} | {k['elementwise']['share'] * 100:.1f} % of the step |")


def sweep(path):
    d = last_line(path)
    cpu = d["cpu_baseline"]["sweep"]
    for key, t in d["sweep"].items():
        print(f"\n**{t['matrix']}** ({t['n']} rows, {t['nnz']} nonzeros), 1 × B200" + (f" vs the reference's MKL kernels on {d['cpu_baseline']['cores']} host threads" if t["matrix"] == cpu["matrix"] else "") + ":\n")
        same = t["matrix"] == cpu["matrix"]
        print("| kernel | B200 ms | B200 GB/s (algorithmic) |" + (" MKL ms | MKL GB/s | speed-up |" if same else ""))
        print("|---|---|---|" + ("---|---|---|" if same else ""))
        for name, v in t["spmv"].items():
            line = f"| SpMV {name} | {v['ms']:.4f} | {v['GBps']:.0f} |"
            if same:
                c = cpu["spmv"].get("mkl_" + name.split("_")[-1]) if name.startswith("packed") else None
                line += f" {c['ms']:.3f} | {c['GBps']:.0f} | {c['ms'] / v['ms']:.0f}× |" if c else " | | |"
            print(line)
        for name, v in t["add_vector"].items():
            line = f"| add_vector {name} (cycle average) | {v['cycle_avg_ms']:.4f} | {v['cycle_avg_GBps']:.0f} |"
            if same:
                c = cpu["add_vector"].get(name)
                line += f" {c['cycle_avg_ms']:.3f} | {c['cycle_avg_GBps']:.0f} | {c['cycle_avg_ms'] / v['cycle_avg_ms']:.0f}× |" if c else " | | |"
            print(line)


if __name__ == "__main__":
    {"scaling": scaling, "sweep": lambda p: sweep(p[0])}[sys.argv[1]](sys.argv[2:])
