#!/usr/bin/env python3
"""profiles/r02_ncu_metrics.json from the `ncu --set full` captures tools/gpu_round2.sh leaves in gpurun_out/:
per workload the kernel durations, the fp64 pipe utilisation time-weighted over the kernels of one step, and the DRAM
bytes of one step (dram__bytes_read.sum + dram__bytes_write.sum, summed over the kernels).  bench.py reads the file for
roofline.fp64_pipe_pct / roofline.traffic.  Usage: python tools/ncu_metrics.py <tag> [edges]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
edges = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
KERNELS = ("k2a_prepare", "k2t_thread_walk", "k2_true_cost")
KEYS = {"ns": "gpu__time_duration.sum", "fp64": "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
        "lts_hit": "lts__t_sector_hit_rate.pct", "l1_hit": "l1tex__t_sector_hit_rate.pct",
        "lanes": "smsp__thread_inst_executed_per_inst_executed.ratio", "regs": "launch__registers_per_thread",
        "warps": "smsp__warps_active.avg.per_cycle_active",
        "long_scoreboard": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "no_instruction": "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"}


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d, u = dict(zip(hdr, vals)), dict(zip(hdr, units))
    res = {}
    for k, name in KEYS.items():
        try:
            v = float(d[name].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = u.get(name, "")
        if k in ("rd", "wr"):  # ncu prints bytes with a scaled unit
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        if k == "ns":
            v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        res[k] = v
    return res


captures = []
for wl in ("c5", "c2"):
    per = {}
    for k in KERNELS:
        rep = os.path.join(ROOT, "gpurun_out", "%s_%s_%s.ncu-rep" % (tag, k, wl))
        if os.path.exists(rep):
            per[k] = raw(rep)
    if len(per) < len(KERNELS):
        continue
    t = sum(v["ns"] for v in per.values())
    captures.append({
        "workload": wl, "edges": edges,
        "source": "ncu --set full --clock-control none of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-plan --no-extra "
                  "--workload %s`, one launch of each kernel (tools/gpu_round2.sh, tools/ncu_metrics.py)" % wl,
        "fp64_pipe_pct": sum(v["fp64"] * v["ns"] for v in per.values()) / t,
        "issue_slots_pct": sum(v["issue"] * v["ns"] for v in per.values()) / t,
        "dram_bytes": sum(v["rd"] + v["wr"] for v in per.values()),
        "step_ns_under_ncu": t,
        "per_kernel": per})
ab = {}
for mode in ("ldg", "smem_tile"):
    rep = os.path.join(ROOT, "gpurun_out", "%s_k2t_tileab_%s.ncu-rep" % (tag, mode))
    if os.path.exists(rep):
        ab[mode] = raw(rep)
out = {"note": "per-launch ncu counters of the kernels of one bench step; durations under ncu are cold-cache and serialised",
       "captures": captures, "map_tile_ab_k2t": ab}
path = os.path.join(ROOT, "profiles", "%s_ncu_metrics.json" % tag)
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1)[:3000])
