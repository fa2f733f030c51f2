#!/bin/bash
mkdir -p gpurun_out
one() { env "$@" timeout 600 python bench.py --workload fit 2>/dev/null | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.readline()); print('$*', round(d['value'],2), 's', int(d['evals_per_s']), 'evals/s', d['objective_evals'])"; }
for rep in 1 2; do
  one GPPLUS_LEAN_LBFGSB=0 GPPLUS_WORKERS_PER_GPU=8
  one GPPLUS_LEAN_LBFGSB=1 GPPLUS_WORKERS_PER_GPU=8
  one GPPLUS_LEAN_LBFGSB=0 GPPLUS_WORKERS_PER_GPU=12
  one GPPLUS_LEAN_LBFGSB=1 GPPLUS_WORKERS_PER_GPU=12
done
nproc; python -c "import os; print(os.cpu_count())"
