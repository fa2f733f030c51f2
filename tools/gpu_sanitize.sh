#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  GPP_GRAPH=0 timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)"
  grep -E "^=========  *(Error|Warning|Race|Invalid|Barrier)" gpurun_out/sanitize_$tool.log | sort | uniq -c | head -8
done
