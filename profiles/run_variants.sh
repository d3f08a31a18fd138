#!/bin/bash
# usage: gpurun -- bash profiles/run_variants.sh "ENV1=a ENV2=b" "ENV1=c" ...   (one quick bench line per variant)
mkdir -p gpurun_out
: > gpurun_out/variants.log
for v in "$@"; do
  echo "== $v" >> gpurun_out/variants.log
  env $v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    ln = ln.strip()
    if ln.startswith('{'):
        d = json.loads(ln); r = d['roofline']
        print('value %.3f G/s  ms/step %.3f  frac %.3f  groups %s  kernel %s' % (d['value'] / 1e9, d['ms_per_step'], r['frac'], r.get('launch_groups'), r['kernel']))
        print('   per-level serial ms: ' + ' '.join('%.3f' % l['ms'] for l in r.get('per_level', [])) + '  scan %.3f' % r.get('knot_scan_ms', 0))
    elif ln: print(ln[:300])
" >> gpurun_out/variants.log
done
cat gpurun_out/variants.log
