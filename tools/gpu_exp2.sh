#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 6 --warmup 3 --no-extras --no-cpu-baseline --no-prior-sweep"
show() { grep '^{' $1 | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); s=d['roofline']['stages_ms']
    print('$2', 'ms/step %.2f' % d['ms_per_step'], 'chol %.2f trtri %.2f lauum %.2f' % (s['cholesky'], s['trtri'], s['lauum']), 'parity', d['parity']['rel_nll'])
"; }
for v in 0 5 10 15; do GPP_STAGGER=$v $B > gpurun_out/exp_st$v.log 2>&1; show gpurun_out/exp_st$v.log stagger$v; done
for v in 0 10; do GPP_PANEL=12 GPP_STAGGER=$v $B > gpurun_out/exp_pb12_st$v.log 2>&1; show gpurun_out/exp_pb12_st$v.log PB12-stagger$v; done
for v in 0 10; do GPP_OVERLAP_INV=0 GPP_STAGGER=$v $B > gpurun_out/exp_noov_st$v.log 2>&1; show gpurun_out/exp_noov_st$v.log no-overlap-stagger$v; done
