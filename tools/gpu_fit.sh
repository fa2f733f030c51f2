#!/bin/bash
mkdir -p gpurun_out
for w in 1 2 4; do
  GPPLUS_WORKERS_PER_GPU=$w timeout 600 python bench.py --workload fit --restarts 15 > gpurun_out/bench_fit_w$w.log 2>&1
  echo "spin workers=$w: $(tail -1 gpurun_out/bench_fit_w$w.log | grep -o '"value[^,]*,') $(tail -1 gpurun_out/bench_fit_w$w.log | grep -o '"objective_evals.*')"
done
for w in 8 16 32 65; do
  GPP_BLOCKING_SYNC=1 GPPLUS_WORKERS_PER_GPU=$w timeout 600 python bench.py --workload fit > gpurun_out/bench_fit_bw$w.log 2>&1
  echo "blocking workers=$w: $(tail -1 gpurun_out/bench_fit_bw$w.log | grep -o '"value[^,]*,') $(tail -1 gpurun_out/bench_fit_bw$w.log | grep -o '"objective_evals.*')"
done
python - <<'PY'
import sys, time
sys.path.insert(0,'.'); sys.path.insert(0,'gp-plus_b200')
import numpy as np, torch, bench
from gpplus_b200.models import GP_Plus
from gpplus_b200.optim.mll_scipy import MLLObjective
Xtr, ytr, Xte, yte, qd = bench._c2_problem()
m = GP_Plus(Xtr, ytr, qual_dict=qd, dtype=torch.float64)
obj = MLLObjective(m, True, [0,0]); obj.enable_fast_path()
th = obj.pack_parameters()
for _ in range(20): obj.fun_fast(th)
t0=time.time()
for k in range(500): obj.fun_fast(th+1e-3*k)
print("single-thread native objective latency us", (time.time()-t0)/500*1e6, m._get_engine().timings())
PY
