"""Digest of ncu output for profiles/: launch-list shares per kernel and the key metrics of `--set full` captures.

    python tools/ncu_digest.py launches gpurun_out/r02_ncu_launches_cd27_256.csv
    python tools/ncu_digest.py full gpurun_out/r02_prof_vpass.ncu-rep [...]      (needs ncu on PATH to read the report)
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def short(name):
    m = re.search(r"(?:<unnamed>::|mpg::)?(\w+)(<[^(]*>)?\(", name)
    return (m.group(1) + (m.group(2) or "")).replace(" ", "") if m else name[:60]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
    hdr = rows[0]
    i_name, i_metric, i_val = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    i_unit = hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[i_metric] != "gpu__time_duration.sum":
            continue
        v = float(r[i_val].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[i_unit], 1e-6)
        base = re.sub(r"<.*", "", short(r[i_name]))
        a = agg.setdefault(base, [0, 0.0])
        a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = [{"kernel": k, "launches": a[0], "ms": round(a[1], 3), "share": round(a[1] / tot, 4)} for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    return {"total_ms": round(tot, 3), "n_launches": sum(a[0] for a in agg.values()), "kernels": out}


def full(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        d = {"kernel": short(r[idx["Kernel Name"]])}
        for k in KEYS:
            if k in idx:
                d[k] = f"{r[idx[k]]} {units[idx[k]]}".strip()
        out.append(d)
    return out


if __name__ == "__main__":
    mode, paths = sys.argv[1], sys.argv[2:]
    if mode == "launches":
        print(json.dumps(launches(paths[0]), indent=1))
    else:
        print(json.dumps({p: full(p) for p in paths}, indent=1))
