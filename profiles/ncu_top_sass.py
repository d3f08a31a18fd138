#!/usr/bin/env python
"""Top stalled SASS instructions of one kernel launch from `ncu -i X.ncu-rep --page source --csv --launch-skip k --launch-count 1`,
with a few instructions of context before each (what the stalled instruction was waiting for).

usage: ncu_top_sass.py source.csv [top N] [context]
"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if r]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hdr = next(r for r in rows if r[0] == "Address")
ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
samp = lambda r: int(r[ci["# Samples"]] or 0)
tot = sum(samp(r) for r in data)
ins = sum(int(r[ci["Instructions Executed"]] or 0) for r in data)
print(f"instructions in kernel {len(data)}, warp instructions executed {ins}, stall samples {tot}")
order = sorted(range(len(data)), key=lambda i: -samp(data[i]))[:topn]
for i in order:
    r = data[i]
    print(f"{samp(r):7d} {100.0 * samp(r) / max(tot, 1):5.1f}%  exec {r[ci['Instructions Executed']]:>10}  {r[1].strip()[:100]}")
    for j in range(max(0, i - ctx), i):
        print(" " * 34 + data[j][1].strip()[:100])
