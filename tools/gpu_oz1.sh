#!/bin/bash
# INT8-sliced GEMM bring-up: correctness checks, then timings (each under its own timeout; a trapped kernel exits non-zero)
mkdir -p gpurun_out
timeout 120 ./tools/oz_lab check > gpurun_out/oz_lab_check.log 2>&1; echo "check rc=$?"; tail -9 gpurun_out/oz_lab_check.log
timeout 240 ./tools/oz_lab perf > gpurun_out/oz_lab_perf.log 2>&1; echo "perf rc=$?"; tail -8 gpurun_out/oz_lab_perf.log
