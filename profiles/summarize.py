#!/usr/bin/env python
"""Turn the ncu captures brought back in gpurun_out/ into the tracked summaries under profiles/.

usage: python profiles/summarize.py <tag>      (expects gpurun_out/{rollout,mpc}_<tag>.ncu-rep, launches_<tag>.csv)
writes profiles/<tag>_summary.json, profiles/<tag>_launches.csv (per-kernel totals) and profiles/traffic.json
"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg",
        "sm__cycles_elapsed.avg.per_second"]
STALLS = "smsp__pcsamp_warps_issue_stalled_"


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {}
        for h, u, v in zip(hdr, units, vals):
            if h in KEYS or h == "Kernel Name":
                d[h] = f"{v} {u}".strip()
            elif h.startswith(STALLS) and "not_issued" not in h and v not in ("0", ""):
                d.setdefault("stall_samples", {})[h[len(STALLS):]] = int(float(v))
        res.append(d)
    return res


def to_bytes(s):
    v, u = s.split()
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


summary = {}
traffic = {}
for name in ("rollout", "mpc", "step"):
    rep = os.path.join(ROOT, "gpurun_out", f"{name}_{tag}.ncu-rep")
    if os.path.exists(rep):
        summary[name] = raw(rep)
        k = summary[name][0]
        traffic["abr_%s_kernel" % name] = to_bytes(k["dram__bytes_read.sum"]) + to_bytes(k["dram__bytes_write.sum"])
lst = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(lst):
    rows = [r for r in csv.reader(open(lst)) if len(r) > 5]
    ix = {h: i for i, h in enumerate(rows[0])}
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        try:
            v = float(r[ix["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        u = r[ix["Metric Unit"]]
        v *= {"ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(u, 1.0)
        tot[r[ix["Kernel Name"]]] += v
        cnt[r[ix["Kernel Name"]]] += 1
    with open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "avg_us", "share_of_gpu_time"])
        allt = sum(tot.values())
        for k in sorted(tot, key=lambda k: -tot[k]):
            w.writerow([k, cnt[k], f"{tot[k]:.1f}", f"{tot[k] / cnt[k]:.2f}", f"{tot[k] / allt:.4f}"])
json.dump(summary, open(os.path.join(ROOT, "profiles", f"{tag}_summary.json"), "w"), indent=1)
json.dump(dict(traffic, tag=tag, unit="bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full)"),
          open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(traffic))
