#!/usr/bin/env python
"""Summarise `ncu --page source --csv` output: instruction mix by opcode and the hottest SASS lines.
usage: ncu_source_summary.py source.csv [kernel_index]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
# split per kernel section
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        sections.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sec = sections[k]
h = {n: i for i, n in enumerate(sec["hdr"])}
print(sec["name"], "instructions:", len(sec["rows"]))
mix, stall = Counter(), Counter()
tot = 0
for r in sec["rows"]:
    op = r[h["Source"]].split()
    if not op:
        continue
    name = op[1] if op[0].startswith("@") else op[0]
    name = name.split(".")[0]
    n = int(r[h["Instructions Executed"]] or 0)
    mix[name] += n
    tot += n
    stall[name] += int(r[h["Warp Stall Sampling (All Samples)"]] or 0)
print("total warp instructions", tot)
for name, n in mix.most_common(28):
    print(f"  {name:10s} {n:12d} {100 * n / tot:5.1f}%   stall samples {stall[name]}")
print("hottest lines by stall samples:")
hot = sorted(sec["rows"], key=lambda r: -int(r[h["Warp Stall Sampling (All Samples)"]] or 0))[:25]
for r in hot:
    print(f"  {r[h['Warp Stall Sampling (All Samples)']]:>7s} exec={r[h['Instructions Executed']]:>10s}  {r[h['Source']].strip()}")
