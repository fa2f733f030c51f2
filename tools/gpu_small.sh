#!/bin/bash
# small-N path: one-launch factorisation (block_potrf_kernel) against the launch chain
mkdir -p gpurun_out
for sb in 0 4 8 16; do
  echo "== GPP_SMALL_BLOCK=$sb"
  GPP_SMALL_BLOCK=$sb timeout 300 python tools/small_eval_probe.py 2>&1 | tail -2
  GPP_SMALL_BLOCK=$sb timeout 600 python bench.py --workload fit 2>&1 | grep '^{' | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('fit_c2 %.3f s  evals %d  evals/s %.0f  best %.6f rrmse %.5f launches %d' % (d['value'], d['objective_evals'], d['evals_per_s'], d['best_neg_log_posterior'], d['test_rrmse'], d.get('gpu_launches_rank0', 0)))
"
done
GPP_SMALL_BLOCK=8 timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_model_gpu.py -m gpu -q -x 2>&1 | tail -2
