#!/bin/bash
# compute-sanitizer memcheck + racecheck over small instances of the hot kernels; logs go to gpurun_out/ (copy the
# summaries into profiles/).  Usage on the GPU box: bash tools/sanitize.sh
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  for c in c2 moment ring relaxed dmd; do
    log=gpurun_out/sanitizer_${tool}_${c}.log
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py $c > $log 2>&1
    echo "$tool $c exit $? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)"
  done
done
