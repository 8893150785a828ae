#!/usr/bin/env python3
"""Summarise one `ncu --set full --import-source on` capture of a kernel: headline counters, stall reasons,
SASS opcode mix and executed instructions aggregated per source function.
Usage: python tools/ncu_summary.py <report.ncu-rep> <edges in the launch> [out.txt]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, n_edges = sys.argv[1], float(sys.argv[2])
out = []


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep] + list(args), capture_output=True, text=True).stdout


rows = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units, d = rows[0], rows[1], dict(zip(rows[0], rows[2]))
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "smsp__warps_active.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum"]
u = dict(zip(hdr, units))
out.append("# %s  (kernel %s)" % (rep.split("/")[-1], d.get("Kernel Name", "")[:60]))
for k in keys:
    out.append("%-72s %s %s" % (k, d.get(k), u.get(k, "")))
out.append("")
out.append("warp stall reasons (cycles per issued instruction):")
for k in hdr:
    if "issue_stalled" in k and "per_issue_active" in k and float(d[k] or 0) > 0.05:
        out.append("  %-24s %.3f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(d[k])))

rows = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "cuda,sass"))))


def funcmap(path):
    m, cur = [], None
    try:
        for line in open(path):
            if line.startswith(("__device__", "__global__", "PPE_HD")):
                mm = re.search(r"(\w+)\s*\(", line)
                if mm:
                    cur = mm.group(1)
            if re.match(r"^k\w+\(", line):
                cur = line.split("(")[0]
            m.append(cur)
    except OSError:
        pass
    return m


fm = {}
cur_file = cur_path = None
h = None
byfn, byop, static = collections.Counter(), collections.Counter(), 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_path, cur_file = r[1], r[1].split("/")[-1]
        if cur_file not in fm:
            fm[cur_file] = funcmap(cur_path)
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        h = r
        ia = h.index("Instructions Executed")
        continue
    if h is None:
        continue
    try:
        n = int(r[ia])
    except ValueError:
        continue
    if r[2] == "-":
        ln = int(r[0])
        name = fm[cur_file][ln - 1] if ln - 1 < len(fm[cur_file]) else None
        byfn["%s:%s" % (cur_file, name) if name else cur_file] += n
    else:
        static += 1
        op = r[3].split()[0] if not r[3].startswith("@") else r[3].split()[1]
        byop[op.split(".")[0]] += n
tot = sum(byop.values())
out.append("")
out.append("executed warp instructions: %d = %.0f per edge; static SASS size %d instructions (%d KB)" % (tot, tot / n_edges, static, static * 16 // 1024))
out.append("opcode mix (per edge): " + ", ".join("%s %.0f" % (k, v / n_edges) for k, v in byop.most_common(14)))
out.append("")
out.append("executed instructions per edge by source function (line-table attribution):")
for k, v in byfn.most_common(22):
    out.append("  %8.1f  %5.1f%%  %s" % (v / n_edges, 100.0 * v / max(1, sum(byfn.values())), k))
text = "\n".join(out) + "\n"
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write(text)
print(text)
