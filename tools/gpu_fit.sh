#!/bin/bash
# 64-restart fit of BASELINE configs[1] with different numbers of restart workers per GPU
mkdir -p gpurun_out
for w in ${@:-4 8 12}; do
  GPPLUS_WORKERS_PER_GPU=$w timeout 600 python bench.py --workload fit 2>/dev/null | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.readline()); print('workers=$w', round(d['value'],2), 's', int(d['evals_per_s']), 'evals/s', d['objective_evals'], 'evals')"
done
