#!/bin/bash
# usage: gpurun --timeout 1500 -- bash profiles/run_sift2d.sh
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sift2d.py tests/test_gpu_spline.py -m gpu -x -q > gpurun_out/pytest_sift2d.log 2>&1
echo "rc=$?" >> gpurun_out/pytest_sift2d.log
tail -30 gpurun_out/pytest_sift2d.log
timeout 400 python profiles/bench_sift2d.py > gpurun_out/bench_sift2d.json 2> gpurun_out/bench_sift2d.err
echo "bench rc=$?"
cat gpurun_out/bench_sift2d.json | cut -c1-1500
tail -5 gpurun_out/bench_sift2d.err
