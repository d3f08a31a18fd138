#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 700 compute-sanitizer --tool $tool --error-exitcode 3 python profiles/sanitize_cases_8f.py > gpurun_out/sanitize8f_$tool.log 2>&1; echo "$tool rc=$?" | tee -a gpurun_out/sanitize8f_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|rc=" gpurun_out/sanitize8f_$tool.log | tail -12
done
