#!/usr/bin/env python
"""Aggregate `ncu --page source --print-source cuda,sass --csv` output per source line.
usage: ncu_lines.py report.csv [top_n]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = defaultdict(lambda: defaultdict(float))
text = {}
h, fname, cur = None, None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        h = None
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        h = r
        continue
    if h is None:
        continue
    if r[0] != "":
        cur = (fname, int(r[0]))
        text[cur] = r[1].strip()
        continue
    if cur is None or len(r) < len(h) or r[2] == "...":
        continue
    for name, v in zip(h[4:], r[4:]):
        try:
            agg[cur][name] += float(v)
        except ValueError:
            pass
tot = sum(a["# Samples"] for a in agg.values())
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
print(f"total samples {tot:.0f}")
gs = defaultdict(float)
for a in agg.values():
    for s in stalls:
        gs[s] += a[s]
print("stall mix:", ", ".join(f"{s[6:]} {100 * v / tot:.1f}%" for s, v in sorted(gs.items(), key=lambda x: -x[1])[:8]))
for key, a in sorted(agg.items(), key=lambda x: -x[1]["# Samples"])[:top]:
    st = sorted(((a[s], s[6:]) for s in stalls), reverse=True)[:3]
    sts = " ".join(f"{n}:{100 * v / max(a['# Samples'], 1):.0f}%" for v, n in st if v > 0)
    print(f"{100 * a['# Samples'] / tot:5.1f}% inst={a['Instructions Executed']:10.0f} {key[0]}:{key[1]:<4d} [{sts}] {text[key][:90]}")
