#!/bin/bash
# usage: gpurun --timeout 1500 -- bash profiles/run_spline.sh
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spline.py -m gpu -x -q > gpurun_out/pytest_spline.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_spline.log
tail -30 gpurun_out/pytest_spline.log
timeout 400 python profiles/bench_spline.py > gpurun_out/bench_spline.jsonl 2> gpurun_out/bench_spline.err
echo "bench rc=$?"
cat gpurun_out/bench_spline.jsonl | cut -c1-900
tail -5 gpurun_out/bench_spline.err
