#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/cfg3_tiles.jsonl
for c in 0 1 2 3; do
  echo "== PYITD_TILE_CFG=$c" >> gpurun_out/cfg3_tiles.jsonl
  PYITD_TILE_CFG=$c timeout 300 python profiles/bench_configs.py --config 3 --dtype f32_mixed >> gpurun_out/cfg3_tiles.jsonl 2>&1
done
cut -c1-330 gpurun_out/cfg3_tiles.jsonl
