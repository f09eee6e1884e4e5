#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_workload.py > $OUT/sanitize_$tool.log 2>&1; echo "rc=$?"
  grep -E "ERROR SUMMARY|SANITIZE_WORKLOAD_OK|RACECHECK SUMMARY|hazard" $OUT/sanitize_$tool.log | head -5
done
