#!/usr/bin/env python
"""Summarise `ncu --page source --csv --print-source cuda,sass` per source line:
warp-instructions executed per line, sorted.  Usage: ncu_source_summary.py rep.ncu-rep [kernel-id]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
per_line = defaultdict(lambda: [0, 0, ""])
cur_file = ""
hdr = None
kernel_seen = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        kernel_seen += 1
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        i_exec = hdr.index("Instructions Executed")
        i_samp = hdr.index("# Samples")
        key = (cur_file, int(r[0]))
        per_line[key][0] += int(r[i_exec] or 0)
        per_line[key][1] += int(r[i_samp] or 0)
        per_line[key][2] = r[1].strip()[:90]
tot = sum(v[0] for v in per_line.values())
print(f"total warp-instructions attributed: {tot}")
import collections
byfile = collections.Counter()
for (f, ln), v in per_line.items():
    byfile[f] += v[0]
print(dict(byfile))
for (f, ln), v in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 45]:
    print(f"{v[0]:10d} {100*v[0]/tot:5.1f}%  samples {v[1]:5d}  {f}:{ln}  {v[2]}")
