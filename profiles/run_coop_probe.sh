#!/bin/bash
# device time of one decomposition of a few signals: cooperative kernel vs look-back launch chain, chunk-size sweep
out=gpurun_out/${1:-coop}; mkdir -p $out
for c in 256 512 768 1024 2048; do echo "chunk $c"; PYITD_COOP_CHUNK=$c python profiles/coop_probe.py --reps 100 2>/dev/null | tee -a $out/coop_chunk_sweep.jsonl | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["coop"]["us_median"], d["lookback"]["us_median"])'; done
for sh in "1 8192" "1 262000" "4 65536" "16 65536" "16 16384" "2 131072"; do set -- $sh; echo "S=$1 n=$2"; python profiles/coop_probe.py --signals $1 --samples $2 --reps 100 2>/dev/null | tee -a $out/coop_shapes.jsonl | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["coop"]["us_median"], d["lookback"]["us_median"])'; done
