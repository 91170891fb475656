#!/usr/bin/env python
"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--clock-control none --csv --log-file X.csv <cmd>`) per kernel: launches, total time, DRAM bytes.
    python scripts/ncu_launch_summary.py X.csv [--units N --unit-name expansion] [--out summary.json]
With --units the totals are also divided by N (e.g. the nodes expanded by the profiled search)."""
import argparse
import csv
import json
import re
from collections import defaultdict

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--units", type=float, default=0)
ap.add_argument("--unit-name", default="unit")
ap.add_argument("--out", default="")
ap.add_argument("--note", default="")
a = ap.parse_args()
rows = []
with open(a.csv, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    rows.append(r)
per = defaultdict(lambda: defaultdict(float))
ids = defaultdict(set)
for r in rows:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("acs::", "")
    v = float(r["Metric Value"].replace(",", "") or 0)
    m, unit = r["Metric Name"], r["Metric Unit"]
    if m == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)  # -> microseconds
    elif m.startswith("dram__bytes"):
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    per[name][m] += v
    ids[name].add(r["ID"])
out = {"source": a.csv, "note": a.note, "kernels": {}}
tot_t = sum(k.get("gpu__time_duration.sum", 0) for k in per.values())
tot_b = 0.0
for name, k in sorted(per.items(), key=lambda kv: -kv[1].get("gpu__time_duration.sum", 0)):
    b = k.get("dram__bytes_read.sum", 0) + k.get("dram__bytes_write.sum", 0)
    tot_b += b
    t = k.get("gpu__time_duration.sum", 0)
    out["kernels"][name] = {"launches": len(ids[name]), "time_us": round(t, 1), "time_share": round(t / tot_t, 4) if tot_t else None,
                            "dram_read_bytes": k.get("dram__bytes_read.sum", 0), "dram_write_bytes": k.get("dram__bytes_write.sum", 0),
                            "dram_GBps": round(b / t / 1e3, 1) if t else None}
out["total_time_us"] = round(tot_t, 1)
out["total_dram_bytes"] = tot_b
if a.units:
    out[f"dram_bytes_per_{a.unit_name}"] = tot_b / a.units
    out[f"{a.unit_name}s"] = a.units
s = json.dumps(out, indent=1)
print(s)
if a.out:
    open(a.out, "w").write(s + "\n")
