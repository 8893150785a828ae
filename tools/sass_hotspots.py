#!/usr/bin/env python3
"""Where a kernel's issue slots and stall samples go, per device function: joins the SASS page of an `ncu --set full
--import-source on` report with the function labels nvdisasm finds in the same kernel's text section of libppe.so.
Usage: python tools/sass_hotspots.py <report.ncu-rep> <kernel-name-fragment> [cubin-name-fragment]"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, frag = sys.argv[1], sys.argv[2]
cub_frag = sys.argv[3] if len(sys.argv) > 3 else "ppe_kernels"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}


def f(r, k):
    try:
        return float(r[ix[k]])
    except (ValueError, IndexError, KeyError):
        return 0.0


tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "path_planner_b200", "libppe.so")], cwd=tmp, capture_output=True)
cubin = [x for x in os.listdir(tmp) if cub_frag in x and x.count("-") == 0][0]
dis = subprocess.run(["nvdisasm", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and frag in l)
idx, labels = 0, []
for l in dis[start + 1:]:
    if l.startswith(".text."):
        break
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        idx += 1
    else:
        m = re.match(r"^\s*\.type\s+(\S+),@function", l)
        if m:
            labels.append((idx, m.group(1)))
if idx != len(data):
    print("warning: %d SASS lines in the report, %d in the cubin (different build?)" % (len(data), idx))
names = subprocess.run(["c++filt"], input="\n".join(re.sub(r"^\$.*\$(_Z)", r"\1", n) for _, n in labels), capture_output=True, text=True).stdout.split("\n")
tot = sum(f(r, "# Samples") for r in data)
totx = sum(f(r, "Instructions Executed") for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("kernel %s: %d SASS instructions, %.3g warp instructions executed, %d samples" % (frag, len(data), totx, tot))
print("stall mix: " + ", ".join("%s %.1f%%" % (s[6:], 100 * v / tot) for s, v in sorted(((s, sum(f(r, s) for r in data)) for s in stalls), key=lambda x: -x[1])[:6]))
labels.append((len(data), "END"))
print("%-13s %8s %8s %8s %8s  %s" % ("sass range", "exec%", "samples%", "no_inst%", "wait%", "function"))
for k in range(len(labels) - 1):
    a, b = labels[k][0], labels[k + 1][0]
    seg = data[a:b]
    s = sum(f(r, "# Samples") for r in seg)
    e = sum(f(r, "Instructions Executed") for r in seg)
    if e == 0:
        continue
    nm = re.sub(r"_INTERNAL_[0-9a-f_]+ppe_\w+_cu_[0-9a-f]+::", "", names[k])
    nm = re.sub(r"\(anonymous namespace\)::", "", nm)
    print("%5d-%-7d %7.1f%% %7.1f%% %7.0f%% %7.0f%%  %s" % (a, b, 100 * e / totx, 100 * s / tot, 100 * sum(f(r, "stall_no_inst") for r in seg) / max(s, 1),
                                                     100 * sum(f(r, "stall_wait") for r in seg) / max(s, 1), nm[:110]))
