#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "blocked_level" > gpurun_out/pytest_blk.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_blk.log
tail -30 gpurun_out/pytest_blk.log
bash profiles/run_variants.sh "PYITD_STREAM_KERNEL=blk8 PYITD_GROUPS=1" "PYITD_STREAM_KERNEL=blk4 PYITD_GROUPS=1" "PYITD_STREAM_KERNEL=blk8 PYITD_GROUPS=2"
