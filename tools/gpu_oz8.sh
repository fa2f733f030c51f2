#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./tools/oz_lab check > gpurun_out/oz_lab_check.log 2>&1; echo "check rc=$?"; tail -12 gpurun_out/oz_lab_check.log
timeout 240 ./tools/oz_lab perf > gpurun_out/oz_lab_perf.log 2>&1; echo "perf rc=$?"; grep -E "^n=.*lauum|K\^-1 shape" gpurun_out/oz_lab_perf.log
timeout 900 python -m pytest tests/test_headline_gpu.py -m gpu -q -x > gpurun_out/oz8_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/oz8_pytest.log
B="python bench.py --steps 4 --warmup 2 --no-extras --no-cpu-baseline --no-prior-sweep --no-dmma-arm"
run() {
  name=$1; shift
  env "$@" timeout 300 $B > gpurun_out/oz8_$name.log 2>&1
  grep '^{' gpurun_out/oz8_$name.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); s=d['roofline']['stages_ms']; print('$name: ms/step %.2f chol %.2f trtri %.2f lauum %.2f  nll %.2e grad %.2e' % (d['ms_per_step'], s['cholesky'], s['trtri'], s['lauum'], d['parity']['rel_nll'], d['parity']['rel_grad']))
" || tail -3 gpurun_out/oz8_$name.log
}
run kinv6 X=1
run kinv7 GPP_OZ_KINV_LEVELS=7
run kinv5 GPP_OZ_KINV_LEVELS=5
