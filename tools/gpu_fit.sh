#!/bin/bash
mkdir -p gpurun_out
for w in 4 8 16 32; do
  GPPLUS_WORKERS_PER_GPU=$w timeout 600 python bench.py --workload fit > gpurun_out/bench_fit_w$w.log 2>&1
  echo "workers=$w: $(tail -1 gpurun_out/bench_fit_w$w.log | cut -c1-700)"
done
GPPLUS_FAST_OBJECTIVE=0 GPPLUS_WORKERS_PER_GPU=8 timeout 600 python bench.py --workload fit > gpurun_out/bench_fit_torchpath.log 2>&1
echo "torch path w=8: $(tail -1 gpurun_out/bench_fit_torchpath.log | cut -c1-700)"
