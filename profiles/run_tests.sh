#!/bin/bash
# usage: gpurun --timeout 1200 -- bash profiles/run_tests.sh "<pytest -k expression>" [sanitize]
mkdir -p gpurun_out
if [ "$2" = "sanitize" ]; then
  timeout 1100 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$1" > gpurun_out/pytest_k.log 2>&1
else
  timeout 900 python -m pytest tests -m gpu -x -q -k "$1" > gpurun_out/pytest_k.log 2>&1
fi
echo "rc=$?" >> gpurun_out/pytest_k.log
tail -60 gpurun_out/pytest_k.log
