#!/bin/bash
# same-box A/B of builds of libpyitd_b200.so: profiles/_ab/*.so (each named on the command line) against the in-tree build
# usage: profiles/run_ab.sh "<a.so> <b.so> ..." [probe args...]
olds=$1; shift
cp pyitd_b200/libpyitd_b200.so /tmp/new.so
pr() { python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(round(d["ms"],4), d.get("fused_pairs"), d.get("stage_ms"))'; }
for rep in 1 2; do
  for old in $olds; do cp $old pyitd_b200/libpyitd_b200.so; echo "$old:"; python profiles/sweep_probe.py "$@" 2>/dev/null | pr; done
  cp /tmp/new.so pyitd_b200/libpyitd_b200.so; echo "new FUSE=0:"; PYITD_SWEEP_FUSE=0 python profiles/sweep_probe.py "$@" 2>/dev/null | pr
  echo "new:"; python profiles/sweep_probe.py "$@" 2>/dev/null | pr
done
