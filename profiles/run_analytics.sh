#!/bin/bash
# usage: gpurun --timeout 1500 -- bash profiles/run_analytics.sh
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_analytics.py -m gpu -x -q > gpurun_out/pytest_analytics.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_analytics.log
tail -30 gpurun_out/pytest_analytics.log
timeout 400 python profiles/bench_analytics.py > gpurun_out/bench_analytics.json 2> gpurun_out/bench_analytics.err
echo "bench rc=$?"
cat gpurun_out/bench_analytics.json | cut -c1-1500
tail -5 gpurun_out/bench_analytics.err
