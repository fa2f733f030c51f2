#!/bin/bash
# N > 16384: segmented K ranges of the INT8-sliced K^-1 against the DMMA arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_headline_gpu.py -m gpu -q -x 2>&1 | tail -2
for n in 20480 32768; do
python - <<PY 2>&1 | tail -3
import sys, time
sys.path[:0] = ['.', 'gp-plus_b200']
import numpy as np
import bench_workloads as W
from gpplus_b200 import _engine as E
n = $n
X, y = W.c4_workload(n)
ys = (y - y.min()) / (y.max() - y.min())
th = W.c4_theta_points(W.c4_model(256))
res = {}
for mode in (1, 0):
    E.set_fp64_mode(mode)
    eng = E.Engine(xq=X, y=ys, kernel=E.KERNEL_MATERN52, n_noise=1, n_mean=1, device=0)
    for k in range(2):
        o = eng.mll_grad(W.c4_natural(th[1]), want_grad=True)
        t = eng.timings()
    res[mode] = (o, t)
    print("n=%d mode %d (reported %d): total %.2f ms chol %.2f trtri %.2f lauum %.2f nll %.12e" % (n, mode, eng.fp64_mode(), t['total'], t['cholesky'], t['trtri'], t['lauum'], o['nll']), flush=True)
    eng.close()
a, b = res[1][0], res[0][0]
ga = np.concatenate([a["d_w"], [a["d_sigma_f2"]], a["d_noise"], a["d_beta"]]); gb = np.concatenate([b["d_w"], [b["d_sigma_f2"]], b["d_noise"], b["d_beta"]])
print("   int8 vs dmma: rel nll %.2e  rel grad %.2e" % (abs(a['nll']-b['nll'])/abs(b['nll']), np.max(np.abs(ga-gb))/np.max(np.abs(gb))))
PY
done
