#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "supplied_knots or single_level or launch_groups" > gpurun_out/pytest_knots.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_knots.log
tail -30 gpurun_out/pytest_knots.log
