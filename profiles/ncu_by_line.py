#!/usr/bin/env python
"""Attribute ncu per-SASS-instruction counters to CUDA source lines.

usage: ncu_by_line.py <source.csv from `ncu --page source --csv`> <kernel substring> <lib.so> [section index] [top N]

The .so must be the build that was profiled (same SASS).  Line info comes from `nvdisasm -g` on the
cubin extracted with `cuobjdump -xelf`; instructions are matched by their offset inside the kernel.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

src_csv, kname, so = sys.argv[1], sys.argv[2], sys.argv[3]
section = int(sys.argv[4]) if len(sys.argv) > 4 else 0
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 40

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout

# offset -> (file, line) for the wanted kernel
rows = list(csv.reader(open(src_csv)))
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        sections.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
sec = [s for s in sections if kname in s["name"]][section]
# mangled-name match: take template args from the demangled name loosely via nvdisasm section order
want = None
line_of = {}
cur_fn, cur_loc = None, None
fn_lines = defaultdict(dict)
for ln in dis.splitlines():
    m = re.match(r"\.text\.(\S+):", ln)
    if m:
        cur_fn, cur_loc = m.group(1), None
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur_loc = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and cur_fn:
        fn_lines[cur_fn][int(m.group(1), 16)] = (cur_loc, m.group(2).strip())
# choose the function whose instruction count and first opcodes match the profiled section
h = {n: i for i, n in enumerate(sec["hdr"])}
addr0 = int(sec["rows"][0][h["Address"]], 16)
nins = len(sec["rows"])
cands = [f for f, d in fn_lines.items() if len(d) == nins]
if not cands:
    cands = sorted(fn_lines, key=lambda f: abs(len(fn_lines[f]) - nins))[:1]
best = None
for f in cands:
    ok = 0
    for r in sec["rows"][:200]:
        off = int(r[h["Address"]], 16) - addr0
        ent = fn_lines[f].get(off)
        if ent and ent[1].split()[0].lstrip("@!P0123456789 ") [:3] == r[h["Source"]].strip().split()[-0 if not r[h["Source"]].strip().startswith("@") else 1][:3]:
            ok += 1
    if best is None or ok > best[0]:
        best = (ok, f)
fn = best[1]
print("kernel:", sec["name"])
print("matched:", fn, f"({len(fn_lines[fn])} vs {nins} instructions)")

by_line = defaultdict(lambda: [0, 0])
tot = 0
for r in sec["rows"]:
    off = int(r[h["Address"]], 16) - addr0
    ent = fn_lines[fn].get(off)
    loc = ent[0] if ent and ent[0] else ("?", 0)
    n = int(r[h["Instructions Executed"]] or 0)
    st = int(r[h["Warp Stall Sampling (All Samples)"]] or 0)
    by_line[loc][0] += n
    by_line[loc][1] += st
    tot += n
stot = sum(v[1] for v in by_line.values())
print(f"total warp instructions {tot}, stall samples {stot}")
srcs = {}
for (f, l), (n, st) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:topn]:
    text = ""
    for root in (os.path.join(os.path.dirname(os.path.abspath(so)), "csrc"), "."):
        pth = os.path.join(root, f)
        if os.path.exists(pth):
            if pth not in srcs:
                srcs[pth] = open(pth).read().splitlines()
            if 0 < l <= len(srcs[pth]):
                text = srcs[pth][l - 1].strip()[:100]
            break
    print(f"{f}:{l:<5d} {n:>11d} {100 * n / tot:5.1f}%  stall {100 * st / max(stot, 1):5.1f}%  {text}")
