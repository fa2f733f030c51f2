#!/bin/bash
# round 2, visit c: lock-step multi-start driver
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_c.log
tail -4 gpurun_out/pytest_gpu_c.log
for i in 1 2; do timeout 600 python bench.py --workload fit > gpurun_out/fit_lockstep$i.log 2>&1; tail -1 gpurun_out/fit_lockstep$i.log | cut -c1-700; done
GPPLUS_LOCKSTEP=0 timeout 600 python bench.py --workload fit > gpurun_out/fit_threads.log 2>&1; tail -1 gpurun_out/fit_threads.log | cut -c1-700
