#!/usr/bin/env python
"""Summarise `ncu --page source --print-source cuda,sass --csv` per CUDA source line: stall samples / instructions.
usage: ncu -i rep.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:<k> | python profiles/src_hot.py [N]"""
import collections
import csv
import sys

N = int(sys.argv[1]) if len(sys.argv) > 1 else 25
agg = collections.defaultdict(lambda: [0, 0, ""])
fname, hdr = "", None
for r in csv.reader(sys.stdin):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        cs, ci = hdr.index("# Samples"), hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) <= max(cs, ci):
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    key = (fname, ln)
    try:
        agg[key][0] += int(r[cs] or 0)
        agg[key][1] += int(r[ci] or 0)
    except ValueError:
        pass
    if r[1].strip():
        agg[key][2] = r[1].strip()[:130]
tot_s = sum(v[0] for v in agg.values()) or 1
tot_i = sum(v[1] for v in agg.values()) or 1
print(f"total samples {tot_s}, total warp instructions {tot_i}")
for (f, ln), (s, i, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:N]:
    print(f"{100*s/tot_s:5.1f}% smp {100*i/tot_i:5.1f}% inst  {f}:{ln:<4} {src}")
