#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_headline_gpu.py -m gpu -q -x > gpurun_out/oz4_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/oz4_pytest.log
B="python bench.py --steps 3 --warmup 1 --no-extras --no-cpu-baseline --no-prior-sweep --no-dmma-arm"
run() {
  name=$1; shift
  env "$@" timeout 300 $B > gpurun_out/oz4_$name.log 2>&1
  grep '^{' gpurun_out/oz4_$name.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); s=d['roofline']['stages_ms']; print('$name: ms/step %.2f chol %.2f trtri %.2f lauum %.2f  nll %.2e grad %.2e' % (d['ms_per_step'], s['cholesky'], s['trtri'], s['lauum'], d['parity']['rel_nll'], d['parity']['rel_grad']))
" || tail -3 gpurun_out/oz4_$name.log
}
run lazy_d3 GPP_TIMELINE=1 GPP_OVERLAP_INV=0 GPP_OZ_STAGGER=3000
grep timeline gpurun_out/oz4_lazy_d3.log | tail -11
run lazy_d3_overlap GPP_OZ_STAGGER=3000
run lazy_d3_overlap6000 X=1
