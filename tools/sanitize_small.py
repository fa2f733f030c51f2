"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'gp-plus_b200'); sys.path.insert(0, 'tests')
import numpy as np
from gpplus_b200 import _engine as E
from problems import engine_kwargs, make_candidates, make_hyper, make_problem
for n, dz in ((200, 2), (700, 0), (1100, 2)):
    p = make_problem(n, 5, 2, dz=dz, n_combo=5 if dz else 0, n_noise=2, seed=3)
    h = make_hyper(p, seed=4)
    eng = E.Engine(**engine_kwargs(p))
    out = eng.mll_grad(h, want_grad=True)
    out = eng.mll_grad(h, want_grad=True)
    c = make_candidates(p, 300)
    mu, var = eng.predict(c["xq"], level_idx=c.get("level_idx"), noise_idx=c.get("noise_idx"), mean_idx=c.get("mean_idx"),
                          include_noise=True)
    print(n, out["nll"], float(mu[0]), float(var[0]))
    eng.close()
