"""End-to-end run of the INT8-sliced path for compute-sanitizer: N = 3200 (25 tiles) with the lazy-panel factorisation
and the block dataflow kernel forced on (GPP_OZ_LAZY_MIN=24 GPP_OZ_LAZY_PB=8), and the same evaluation on DMMA."""
import os
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'gp-plus_b200'); sys.path.insert(0, 'tests')
os.environ.setdefault("GPP_OZ_LAZY_MIN", "24")
os.environ.setdefault("GPP_OZ_LAZY_PB", "8")
import numpy as np
from gpplus_b200 import _engine as E
from problems import engine_kwargs, make_hyper, make_problem
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3200
p = make_problem(n, 5, 2, dz=2, n_combo=5, n_noise=2, seed=3)
h = make_hyper(p, seed=4)
res = {}
for mode in (E.FP64_INT8, E.FP64_DMMA):
    E.set_fp64_mode(mode)
    eng = E.Engine(**engine_kwargs(p))
    out = eng.mll_grad(h, want_grad=True)
    res[mode] = out
    print("mode", eng.fp64_mode(), out["nll"], flush=True)
    eng.close()
a, b = res[E.FP64_INT8], res[E.FP64_DMMA]
print("rel nll", abs(a["nll"] - b["nll"]) / abs(b["nll"]), "max |d_w| diff", float(np.max(np.abs(a["d_w"] - b["d_w"]))))
