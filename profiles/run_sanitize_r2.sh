#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 700 compute-sanitizer --tool $tool --error-exitcode 3 python profiles/sanitize_cases_r2.py > gpurun_out/sanitize_r2_$tool.log 2>&1; echo "$tool rc=$?" | tee -a gpurun_out/sanitize_r2_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok|rc=|Error|hazard" gpurun_out/sanitize_r2_$tool.log | sort | uniq -c | sort -rn | head -14
done
