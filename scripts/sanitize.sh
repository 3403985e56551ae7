#!/bin/bash
# compute-sanitizer over every kernel at small sizes (scripts/sanitizer_driver.py); summaries under gpurun_out/ (copy into profiles/).
# usage: scripts/sanitize.sh <tag>
tag=${1:-r2}
out=gpurun_out; mkdir -p $out
for tool in memcheck racecheck synccheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitizer_driver.py > $out/${tag}_sanitizer_${tool}.log 2>&1
  echo "== $tool: rc $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitizer driver ran|Error|hazard" $out/${tag}_sanitizer_${tool}.log | head -8
done
