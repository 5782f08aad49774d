#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  for w in ba match; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_small.py $w > gpurun_out/r2_sanitize_${tool}_$w.log 2>&1
    echo "$tool $w: exit $? | $(grep -c 'ERROR SUMMARY' gpurun_out/r2_sanitize_${tool}_$w.log) | $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/r2_sanitize_${tool}_$w.log | tail -1)"
  done
done
