#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  GPP_GRAPH=0 timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool (streams): $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)"
done
GPP_GRAPH=1 timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/sanitize_memcheck_graph.log 2>&1
echo "== memcheck (graph replay): $(grep -E 'ERROR SUMMARY' gpurun_out/sanitize_memcheck_graph.log | tail -1)"
grep -c "^[0-9]" gpurun_out/sanitize_memcheck_graph.log
