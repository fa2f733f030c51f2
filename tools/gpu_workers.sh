#!/bin/bash
for n in 1536 3072 6144; do for w in 1 2 4 8; do
  GPPLUS_WORKERS_PER_GPU=$w timeout 600 python bench.py --workload fit --fit-config c4 --size $n --restarts 15 --maxiter 12 2>/dev/null | tail -1 | python -c "
import sys, json
d=json.loads(sys.stdin.readline()); print('n=$n workers=$w', round(d['value'],2), 's', int(d['evals_per_s']), 'evals/s', d['objective_evals'])"
done; done
