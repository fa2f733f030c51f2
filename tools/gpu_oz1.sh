#!/bin/bash
# INT8-sliced GEMM bring-up: correctness checks, then timings (each under its own timeout; a trapped kernel exits non-zero)
mkdir -p gpurun_out
for v in oz_lab oz_lab_bk64; do
  [ -x tools/$v ] || continue
  timeout 120 ./tools/$v check > gpurun_out/${v}_check.log 2>&1; echo "$v check rc=$?"; tail -9 gpurun_out/${v}_check.log
  timeout 240 ./tools/$v perf > gpurun_out/${v}_perf.log 2>&1; echo "$v perf rc=$?"; tail -8 gpurun_out/${v}_perf.log
done
