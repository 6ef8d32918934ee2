#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  echo "== $tool"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 20 python scripts/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1; echo "rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_small done" gpurun_out/sanitize_$tool.log | tail -3
  grep -E "Invalid|Race reported|hazard" gpurun_out/sanitize_$tool.log | head -10
done
