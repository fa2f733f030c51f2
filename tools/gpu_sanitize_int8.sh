#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/sanitize_int8.py 2>&1 | tail -3
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_int8.py > gpurun_out/sanitize_int8_$tool.log 2>&1
  echo "== $tool: rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_int8_$tool.log | tail -1)"
  grep -E "^mode|rel nll" gpurun_out/sanitize_int8_$tool.log | tail -3
done
