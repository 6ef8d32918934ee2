#!/bin/bash
# compute-sanitizer over scripts/sanitize_small.py: memcheck on every variant, racecheck (shared-memory hazards) on the defaults
mkdir -p gpurun_out
echo "== memcheck"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python scripts/sanitize_small.py > gpurun_out/r2_sanitize_memcheck.log 2>&1; echo "rc=$?"
grep -E "ERROR SUMMARY|sanitize_small done" gpurun_out/r2_sanitize_memcheck.log | tail -3
grep -E "Invalid|out of bounds" gpurun_out/r2_sanitize_memcheck.log | head -10
echo "== racecheck"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python scripts/sanitize_small.py --light > gpurun_out/r2_sanitize_racecheck.log 2>&1; echo "rc=$?"
grep -E "RACECHECK SUMMARY|sanitize_small done" gpurun_out/r2_sanitize_racecheck.log | tail -3
grep -E "Race reported|hazard" gpurun_out/r2_sanitize_racecheck.log | head -10
