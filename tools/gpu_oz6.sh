#!/bin/bash
# diagonal-block dataflow kernel: parity + timeline + comparison with the launch-per-step chain
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_headline_gpu.py tests/test_engine_gpu.py -m gpu -q -x > gpurun_out/oz6_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/oz6_pytest.log
B="python bench.py --steps 3 --warmup 1 --no-extras --no-cpu-baseline --no-prior-sweep --no-dmma-arm"
run() {
  name=$1; shift
  env "$@" timeout 300 $B > gpurun_out/oz6_$name.log 2>&1
  grep '^{' gpurun_out/oz6_$name.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); s=d['roofline']['stages_ms']; print('$name: ms/step %.2f chol %.2f trtri %.2f lauum %.2f  nll %.2e grad %.2e' % (d['ms_per_step'], s['cholesky'], s['trtri'], s['lauum'], d['parity']['rel_nll'], d['parity']['rel_grad']))
" || tail -3 gpurun_out/oz6_$name.log
}
run blk24_noverlap GPP_TIMELINE=1 GPP_OVERLAP_INV=0
grep timeline gpurun_out/oz6_blk24_noverlap.log | tail -11
run blk0_noverlap GPP_OVERLAP_INV=0 GPP_BLOCK_CTAS=0
run blk12_noverlap GPP_OVERLAP_INV=0 GPP_BLOCK_CTAS=12
run blk48_noverlap GPP_OVERLAP_INV=0 GPP_BLOCK_CTAS=48
run blk24 X=1
run blk0 GPP_BLOCK_CTAS=0
