#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_d.log
tail -4 gpurun_out/pytest_gpu_e.log
for i in 1 2 3; do GPPLUS_LOCKSTEP_PROFILE=1 timeout 600 python bench.py --workload fit > gpurun_out/fit_lockstep_e$i.log 2>&1; grep lockstep gpurun_out/fit_lockstep_e$i.log; tail -1 gpurun_out/fit_lockstep_e$i.log | cut -c1-420; done
