#!/bin/bash
# INT8-sliced path inside the engine: headline parity (N = 4096 / 8192 / 16384 against the oracle) and stage times
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_headline_gpu.py -m gpu -q -x > gpurun_out/oz2_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/oz2_pytest.log
B="python bench.py --steps 6 --warmup 3 --no-extras --no-cpu-baseline --no-prior-sweep"
for mode in int8 dmma; do
GPP_FP64=$mode timeout 600 $B > gpurun_out/oz2_bench_$mode.log 2>&1
grep '^{' gpurun_out/oz2_bench_$mode.log | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('$mode ms/step %.2f' % d['ms_per_step'], d['roofline']['stages_ms'], d['parity'])
" || tail -5 gpurun_out/oz2_bench_$mode.log
done
