#!/bin/bash
# INT8-sliced GEMM bring-up: correctness checks, then timings (each under its own timeout; a trapped kernel exits non-zero)
mkdir -p gpurun_out
for v in oz_lab; do
  [ -x tools/$v ] || continue
  timeout 120 ./tools/$v check > gpurun_out/${v}_check.log 2>&1; echo "$v check rc=$?"; tail -2 gpurun_out/${v}_check.log
  timeout 240 ./tools/$v perf > gpurun_out/${v}_perf.log 2>&1; echo "$v perf rc=$?"; grep "^n=" gpurun_out/${v}_perf.log
done
