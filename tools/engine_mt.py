"""Throughput of concurrent engine handles from W host threads (ctypes releases the GIL inside the call)."""
import sys, time, threading, os
sys.path.insert(0, '.'); sys.path.insert(0, 'gp-plus_b200')
import numpy as np, torch, bench, copy
from gpplus_b200.models import GP_Plus
from gpplus_b200.optim.mll_scipy import MLLObjective

Xtr, ytr, Xte, yte, qd = bench._c2_problem()
base = GP_Plus(Xtr, ytr, qual_dict=qd, dtype=torch.float64)
obj0 = MLLObjective(base, True, [0, 0]); obj0.enable_fast_path()
th0 = obj0.pack_parameters()

def run(W, iters=1500):
    objs = []
    for w in range(W):
        m = GP_Plus(Xtr.clone(), ytr.clone(), qual_dict=qd, dtype=torch.float64)
        o = MLLObjective(m, True, [0, 0])
        assert o.enable_fast_path()
        objs.append(o)
        o.fun_fast(th0)  # create engine + layout + graph
    start = threading.Barrier(W + 1)
    def work(o):
        eng = o.model._get_engine()
        th = th0.copy()
        start.wait()
        for k in range(iters):
            th[3] = 1e-4 * k
            eng.objective(th, True)
    ts = [threading.Thread(target=work, args=(o,)) for o in objs]
    for t in ts: t.start()
    start.wait(); t0 = time.time()
    for t in ts: t.join()
    dt = time.time() - t0
    for o in objs: o.model.release_engine()
    return W * iters / dt

for W in (1, 2, 4, 8, 16):
    print("threads %2d: %.0f evals/s" % (W, run(W)), flush=True)
