#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest rc=$rc" | tee -a gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ]; then exit 0; fi
bash profiles/run_variants.sh "PYITD_GROUPS=1" "PYITD_GROUPS=2"
timeout 300 python profiles/bench_configs.py --config 3 --dtype f32_mixed 2>&1 | cut -c1-200
timeout 300 python profiles/bench_configs.py --config 4 --dtype f32_mixed 2>&1 | cut -c1-200
