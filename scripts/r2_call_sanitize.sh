#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_small.py > gpurun_out/r2j_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|PMM_DF|nni " gpurun_out/r2j_sanitizer_$tool.log | tail -8
done
