#!/bin/bash
# compute-sanitizer over the smallest run that touches every kernel: memcheck, racecheck, initcheck (logs under gpurun_out/)
TAG=${1:-sanitize}
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > gpurun_out/${TAG}_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_small done" gpurun_out/${TAG}_$tool.log | tail -3
done
