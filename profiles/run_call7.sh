#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "both_level_kernels" > gpurun_out/pytest_quick.log 2>&1; rc=$?; echo "quick pytest rc=$rc"
tail -5 gpurun_out/pytest_quick.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
bash profiles/run_variants.sh "PYITD_GROUPS=1" "PYITD_GROUPS=2" "PYITD_GROUPS=4"
