#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "strided" > gpurun_out/pytest_strided.log 2>&1; rc=$?; echo "strided pytest rc=$rc"
tail -25 gpurun_out/pytest_strided.log | cut -c1-200
if [ $rc -ne 0 ]; then exit 0; fi
: > gpurun_out/cfg3_strided.jsonl
timeout 300 python profiles/bench_configs.py --config 3 --dtype f32_mixed >> gpurun_out/cfg3_strided.jsonl 2>&1
timeout 300 python profiles/bench_configs.py --config 3 --dtype f32 >> gpurun_out/cfg3_strided.jsonl 2>&1
cut -c1-330 gpurun_out/cfg3_strided.jsonl
