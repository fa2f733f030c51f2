#!/bin/bash
run() { env "$@" timeout 600 python bench.py --steps 8 --warmup 3 $EXTRA 2>gpurun_out/ab_err.log | python -c "
import sys, json
d=json.loads(sys.stdin.readline()); print('$* $EXTRA', round(d['ms_per_step'],2), 'ms', {k: round(v,2) for k,v in d['roofline']['stages_ms'].items() if k in ('cholesky','trtri','lauum')})"; }
run GPP_OVERLAP_INV=1
run GPP_OVERLAP_INV=2
run GPP_OVERLAP_INV=0
run GPP_OVERLAP_INV=1 GPP_PANEL=4
EXTRA="--n 8192"
run GPP_OVERLAP_INV=1
run GPP_OVERLAP_INV=2
run GPP_OVERLAP_INV=0
EXTRA="--n 2048"
run GPP_OVERLAP_INV=1
run GPP_OVERLAP_INV=0
