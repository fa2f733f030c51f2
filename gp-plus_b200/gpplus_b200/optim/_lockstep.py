"""Lock-step multi-start L-BFGS-B: every in-flight restart of a GPU is driven by ONE host thread.

The reference fans the ``num_restarts + 1`` L-BFGS-B runs out over joblib/loky processes
(optim/mll_scipy.py:287-293).  For small training sets (N <= 2048) one objective evaluation on the B200 is a
latency-bound chain of ~25 tiny kernels that leaves most of the GPU idle, so throughput comes from keeping many
restarts in flight.  Host threads (one per restart) serialise on the interpreter lock once there are more than ~8
of them; this driver instead runs scipy's compiled L-BFGS-B routine (``scipy.optimize._lbfgsb.setulb``) in its
reverse-communication form -- exactly the loop of ``scipy.optimize._lbfgsb_py._minimize_lbfgsb`` -- for up to 64
restarts at once from a single thread:

    ring of slots:  collect slot k (waits for ITS evaluation only) -> feed (f, g) to its L-BFGS-B state -> advance
                    until the next (f, g) request -> enqueue that evaluation (gpp_objective_enqueue, no waiting)

While the host handles slot k the evaluations of all other slots run on the GPU (one engine handle and stream per
slot).  Each restart's iterates are those of ``scipy.optimize.minimize(method="L-BFGS-B")`` from the same start --
the state machines are independent -- so results do not depend on the number of slots or on which rank ran them.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
from scipy.optimize import OptimizeResult
from scipy.optimize import _lbfgsb
from scipy.optimize._lbfgsb_py import LbfgsInvHessProduct, status_messages, task_messages

from .._engine import NanError, NotPSDError

try:
    from scipy._lib._util import HAS_ILP64 as _ILP64  # scipy >= 1.15 keeps it here
except Exception:  # pragma: no cover
    _ILP64 = False
_INT = np.int64 if _ILP64 else np.int32


class _Chain:
    """One L-BFGS-B run as a resumable state machine (same arrays and bookkeeping as _minimize_lbfgsb)."""

    def __init__(self, x0, lo, hi, options):
        self.m = int(options.get("maxcor", 10))
        self.maxls = int(options.get("maxls", 20))
        self.maxiter = int(options.get("maxiter", 15000))
        self.maxfun = int(options.get("maxfun", 15000))
        self.factr = float(options.get("ftol", 2.2204460492503131e-09)) / np.finfo(float).eps
        self.pgtol = float(options.get("gtol", 1e-5))
        x0 = np.asarray(x0, dtype=np.float64).ravel()
        n = x0.shape[0]
        self.nbd = np.zeros(n, dtype=_INT)
        self.low = np.zeros(n, dtype=np.float64)
        self.up = np.zeros(n, dtype=np.float64)
        if lo is not None:
            x0 = np.clip(x0, lo, hi)
            for i in range(n):
                has_l, has_u = np.isfinite(lo[i]), np.isfinite(hi[i])
                if has_l:
                    self.low[i] = lo[i]
                if has_u:
                    self.up[i] = hi[i]
                self.nbd[i] = {(False, False): 0, (True, False): 1, (True, True): 2, (False, True): 3}[(has_l, has_u)]
        m = self.m
        self.x = np.array(x0, dtype=np.float64)
        self.f = np.array(0.0, dtype=np.float64)
        self.g = np.zeros(n, dtype=np.float64)
        self.wa = np.zeros(2 * m * n + 5 * n + 11 * m * m + 8 * m, np.float64)
        self.iwa = np.zeros(3 * n, dtype=_INT)
        self.task = np.zeros(2, dtype=_INT)
        self.ln_task = np.zeros(2, dtype=_INT)
        self.lsave = np.zeros(4, dtype=_INT)
        self.isave = np.zeros(44, dtype=_INT)
        self.dsave = np.zeros(29, dtype=np.float64)
        self.nit = 0
        self.nfev = 0
        self._x_eval = None   # last point evaluated (scipy's ScalarFunction does not re-evaluate an unchanged x)
        self._f_eval = 0.0
        self._g_eval = None

    def advance(self) -> bool:
        """Run the optimiser until it needs f and g at ``self.x`` (True) or has finished (False)."""
        while True:
            _lbfgsb.setulb(self.m, self.x, self.low, self.up, self.nbd, self.f, self.g, self.factr, self.pgtol,
                           self.wa, self.iwa, self.task, self.lsave, self.isave, self.dsave, self.maxls, self.ln_task)
            t = self.task[0]
            if t == 3:
                if self._x_eval is not None and np.array_equal(self.x, self._x_eval):
                    # same point requested again: hand back the cached value and gradient, no new evaluation
                    self.f[()] = self._f_eval
                    self.g[:] = self._g_eval
                    continue
                return True
            if t == 1:
                self.nit += 1
                if self.nit >= self.maxiter:
                    self.task[0], self.task[1] = 5, 504
                elif self.nfev > self.maxfun:
                    self.task[0], self.task[1] = 5, 502
            else:
                return False

    def feed(self, value: float):
        """f at ``self.x``; the gradient was written into ``self.g`` by the engine."""
        self.f[()] = value
        self.nfev += 1
        self._x_eval = self.x.copy()
        self._f_eval = value
        self._g_eval = self.g.copy()

    def result(self) -> OptimizeResult:
        if self.task[0] == 4:
            warnflag = 0
        elif self.nfev > self.maxfun or self.nit >= self.maxiter:
            warnflag = 1
        else:
            warnflag = 2
        m, n = self.m, self.x.shape[0]
        s = self.wa[0: m * n].reshape(m, n)
        y = self.wa[m * n: 2 * m * n].reshape(m, n)
        n_corrs = min(int(self.isave[30]), m)
        msg = status_messages[int(self.task[0])] + ": " + task_messages[int(self.task[1])]
        return OptimizeResult(fun=float(self.f), jac=self.g.copy(), nfev=self.nfev, njev=self.nfev, nit=self.nit,
                              status=warnflag, message=msg, x=self.x.copy(), success=(warnflag == 0),
                              hess_inv=LbfgsInvHessProduct(s[:n_corrs].copy(), y[:n_corrs].copy()))


def slots_for(n_train: int) -> int:
    """Restarts kept in flight per GPU: the GPU's throughput on these latency-bound evaluations saturates around
    16-32 concurrent streams; more slots only shorten the tail (every restart starts in the first wave)."""
    if n_train <= 1024:
        return 64
    return 32


def run_lockstep(likobj, theta0_list: List[np.ndarray], work, options: Dict, lo: Optional[np.ndarray],
                 hi: Optional[np.ndarray], device: int, max_slots: Optional[int] = None) -> Dict[int, object]:
    """Drive restarts claimed from ``work`` on GPU ``device`` until the queue is drained; returns {index: result}
    with numerical failures (NotPSDError / NanError) stored as values, like ``_fit_model_from_state``."""
    model = likobj.model
    spec = likobj._fast.layout_spec()
    n_train = int(model.train_targets.shape[0])
    n_slots = max(1, min(max_slots or slots_for(n_train), work.count))
    results: Dict[int, object] = {}
    slots = []  # [engine, chain, restart index]
    import os
    import time
    profile = os.environ.get("GPPLUS_LOCKSTEP_PROFILE", "0") != "0"
    t_begin = time.time()
    t_setup = 0.0

    def start(slot) -> bool:
        """Claim restarts until one needs an evaluation; enqueue it.  False when the queue is drained."""
        while True:
            i = work.claim()
            if i is None:
                slot[1], slot[2] = None, None
                return False
            chain = _Chain(theta0_list[i], lo, hi, options)
            if chain.advance():
                slot[1], slot[2] = chain, i
                slot[0].objective_enqueue(chain.x, True)
                return True
            results[i] = chain.result()

    try:
        # every handle first (with the GPU idle), then the first evaluation of every slot
        t0 = time.time()
        kwargs = model._engine_kwargs(device)
        for _ in range(n_slots):
            eng = model._new_engine(device, kwargs)
            eng.set_theta_layout(spec)
            slots.append([eng, None, None])
        t_setup = time.time() - t0
        for slot in slots:
            if not start(slot):
                break
        active = [s for s in slots if s[1] is not None]
        while active:
            for slot in active:
                eng, chain, i = slot
                try:
                    value = eng.objective_collect(chain.g)
                except (NotPSDError, NanError) as e:
                    results[i] = e  # unstable point: the restart is scored +inf by the caller
                    start(slot)
                    continue
                chain.feed(value)
                if chain.advance():
                    eng.objective_enqueue(chain.x, True)
                else:
                    results[i] = chain.result()
                    start(slot)
            active = [s for s in active if s[1] is not None]
    finally:
        t_loop = time.time()
        for slot in slots:
            slot[0].close()
        if profile:
            print("[lockstep] device %d: %d slots, %d restarts, engine setup %.3f s, total before close %.3f s, "
                  "close %.3f s" % (device, len(slots), len(results), t_setup, t_loop - t_begin,
                                    time.time() - t_loop), flush=True)
    return results
