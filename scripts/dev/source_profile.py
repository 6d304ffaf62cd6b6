"""Aggregates an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line (development aid)."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
kern = None
fname = ''
hdr = None
agg = {}
cur_line = None
with open(path, newline="") as f:
    for row in csv.reader(f):
        if not row:
            continue
        if row[0] == "File Path":
            fname = row[1].split("/")[-1]
            continue
        if row[0] == "Function Name":
            kern = row[1].split("(")[0].replace("void nlz::", "")
            agg.setdefault(kern, defaultdict(lambda: [0, 0, 0, ""]))
            hdr = None
            cur_line = None
            continue
        if row[0] == "Line No":
            hdr = row
            iS = hdr.index("# Samples"); iI = hdr.index("Instructions Executed"); iT = hdr.index("Thread Instructions Executed")
            continue
        if hdr is None or kern is None or len(row) < len(hdr) - 5:
            continue
        if row[0].strip():                       # a CUDA source line starts a group; SASS rows follow with empty "Line No"
            cur_line = (fname + ":" + row[0].strip(), row[1].strip()[:110])
            continue
        if cur_line is None:
            continue
        try:
            s = int(row[iS] or 0); i = int(row[iI] or 0); t = int(row[iT] or 0)
        except ValueError:
            continue
        a = agg[kern][cur_line[0]]
        a[0] += s; a[1] += i; a[2] += t; a[3] = cur_line[1]
for k, lines in agg.items():
    ts = sum(v[0] for v in lines.values()); ti = sum(v[1] for v in lines.values())
    print(f"== {k}: samples {ts}, warp instructions {ti}")
    for ln, v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"  {ln:>16} samples {100 * v[0] / max(ts, 1):5.1f}%  inst {100 * v[1] / max(ti, 1):5.1f}%  thr/inst {v[2] / max(v[1], 1):5.1f}  {v[3]}")
