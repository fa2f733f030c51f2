#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 4 --warmup 2 --no-extras --no-cpu-baseline --no-prior-sweep --no-dmma-arm"
run() {
  name=$1; shift
  env "$@" timeout 300 $B > gpurun_out/oz7_$name.log 2>&1
  grep '^{' gpurun_out/oz7_$name.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); s=d['roofline']['stages_ms']; print('$name: ms/step %.2f chol %.2f trtri %.2f lauum %.2f  nll %.2e grad %.2e' % (d['ms_per_step'], s['cholesky'], s['trtri'], s['lauum'], d['parity']['rel_nll'], d['parity']['rel_grad']))
" || tail -3 gpurun_out/oz7_$name.log
}
run pb12 X=1
run pb8 GPP_OZ_LAZY_PB=8
run pb10 GPP_OZ_LAZY_PB=10
run pb16 GPP_OZ_LAZY_PB=16
run pb6 GPP_OZ_LAZY_PB=6
