"""One MLL+gradient evaluation of the C2 model (n=500, 2 categorical x 5 levels) with CUDA-graph replay off, so that
ncu lists the kernels of the latency-bound small-N path one by one.  Usage: GPP_GRAPH=0 ncu ... python tools/small_eval_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gp-plus_b200"))
import time  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from gpplus_b200 import _engine as E  # noqa: E402
from gpplus_b200.models import GP_Plus  # noqa: E402
from gpplus_b200.optim.mll_scipy import MLLObjective  # noqa: E402

Xtr, ytr, Xte, yte, qd = bench._c2_problem()
m = GP_Plus(Xtr, ytr, qual_dict=qd, dtype=torch.float64)
obj = MLLObjective(m, True, [0, 0])
assert obj.enable_fast_path()
th = obj.pack_parameters() + 0.05
l0 = E.launch_count()
for k in range(3):
    t0 = time.time()
    f, g = obj.fun_fast(th + 0.01 * k)
    dt = time.time() - t0
    print("eval %d: %.3f ms, launches so far %d" % (k, 1e3 * dt, E.launch_count() - l0), flush=True)
