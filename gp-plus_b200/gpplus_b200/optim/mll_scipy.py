"""Multi-start MAP fitting with scipy optimisers over the B200 engine.

Mirrors optim/mll_scipy.py of the reference: ``marginal_log_likelihood`` (:37-60), ``MLLObjective``
(:63-127: theta packing in ``named_parameters`` order, float32 cast of theta, ``fun`` returning
``(nll, grad)``), ``_sample_from_prior`` (:130-138), ``get_bounds`` (:149-183),
``_fit_model_from_state`` (:187-240, numerical failures are *returned* as values) and
``fit_model_scipy`` (:244-307, ``num_restarts + 1`` prior draws, argmin over restarts, best theta
loaded into the model, returns ``(results, best_nll)``).

What changed underneath: the objective is one ``gpp_mll_grad`` call (fused covariance build, blocked
DMMA Cholesky, triangular inverse, fused gradient reduction) instead of a dense torch forward +
autograd backward, and the joblib/loky process fan-out (:287-293) becomes a restart scheduler: host
threads that each own a model copy and an engine handle, spread over the visible GPUs; under
``torchrun`` the restarts are claimed from a cross-rank work queue (``parallel.RestartQueue``; one GPU per
rank) and only ``(nll, theta)`` of every restart is gathered (``parallel.gather_restarts``).
"""
from __future__ import annotations

import copy
import os
import queue
import threading
from collections import OrderedDict
from functools import reduce
from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import torch
from scipy.optimize import Bounds, OptimizeResult, minimize

from .. import _engine, parallel
from .._compat import NanError, NotPSDError, settings as gptsettings
from ..utils.interval_score import interval_score_function

# theta is cast to this dtype on every evaluation, exactly like the reference (mll_scipy.py:32-35,97);
# set ``tkwargs["dtype"] = torch.float64`` to keep the optimiser's iterates unrounded.
tkwargs = {
    "dtype": torch.float,
    "device": torch.device("cpu"),
}

_REG_NAMES = ["fci", "h1", "h2", "h3", "h4", "h5", "h6", "h8", "h9", "h10", "h11", "h12", "fce"]


def marginal_log_likelihood(model, add_prior: bool, regularization_parameter=[0, 0]):
    """log p(y | X, theta) (+ log-priors, - weight penalties): a differentiable float64 torch scalar whose
    data term and gradient come from the GPU engine."""
    out = model.log_marginal()
    if add_prior:
        for _, module, prior, closure, _ in model.named_priors():
            out = out + prior.log_prob(closure(module)).sum()
    l1 = 0
    l2 = 0
    for name, param in model.named_parameters():
        if name in _REG_NAMES or name in ["nn_model." + s + ".bias" for s in _REG_NAMES]:
            l2 = l2 + torch.norm(param)
            l1 = l1 + torch.sum(torch.abs(param))
    out = out - (regularization_parameter[0] * l1 + regularization_parameter[1] * l2)
    if getattr(model, "interval_score", False) is True:
        # mll_scipy.py:57-59: ``output = model(*model.train_inputs)`` is evaluated in TRAINING mode, i.e. it is the
        # PRIOR at the training inputs: mean m(x_i) (the constant of the point's mean group) and variance
        # diag(Sigma) = sigma_f^2 (clamped at settings.min_variance).  Both are differentiable functions of the raw
        # parameters, so the penalty contributes to the gradient through the output scale and the mean constants.
        mean, var = model.prior_mean_and_variance()
        sd = var.clamp_min(gptsettings.min_variance).sqrt()
        score, _ = interval_score_function(mean + 1.96 * sd, mean - 1.96 * sd, model.y_scaled)
        return out - 0.08 * torch.abs(out) * score
    return out


class MLLObjective:
    """theta <-> model parameters, and ``fun(theta) -> (nll, d nll / d theta)`` for scipy."""

    def __init__(self, model, add_prior, regularization_parameter):
        self.model = model
        self.add_prior = add_prior
        self.regularization_parameter = regularization_parameter
        self.param_shapes = OrderedDict()
        self._fast = None
        self._native = True
        for n, p in self.model.named_parameters():
            if p.requires_grad:
                self.param_shapes[n] = p.size() if len(p.size()) > 0 else torch.Size([1])

    def pack_parameters(self) -> np.ndarray:
        parts = [p.cpu().data.numpy().ravel() for _, p in self.model.named_parameters() if p.requires_grad]
        return np.concatenate(parts)

    def unpack_parameters(self, x: np.ndarray) -> "OrderedDict[str, torch.Tensor]":
        i = 0
        named = OrderedDict()
        for n, shape in self.param_shapes.items():
            length = reduce(lambda a, b: a * b, shape)
            named[n] = torch.from_numpy(np.asarray(x[i:i + length]).reshape(*shape)).to(**tkwargs)
            i += length
        return named

    def pack_grads(self) -> np.ndarray:
        grads = []
        for _, p in self.model.named_parameters():
            if p.requires_grad:
                g = p.grad if p.grad is not None else torch.zeros_like(p)
                grads.append(g.cpu().data.numpy().ravel())
        return np.concatenate(grads).astype(np.float64)

    def _load(self, x: np.ndarray):
        params = dict(self.model.named_parameters())
        with torch.no_grad():
            for n, v in self.unpack_parameters(x).items():
                params[n].copy_(v.reshape(params[n].shape))

    def enable_fast_path(self) -> bool:
        """Compile the closed-form host path (optim/_fast_objective.py) for this model and validate it against
        the torch path; returns False (and keeps the torch path) when the model is outside its closed forms."""
        from . import _fast_objective as FO
        self._fast = None
        if os.environ.get("GPPLUS_FAST_OBJECTIVE", "1") == "0":
            return False
        fast = FO.build(self.model, self.add_prior, self.regularization_parameter)
        if fast is not None and FO.self_check(self, fast):
            self._fast = fast
        # the closed forms evaluated by numpy on the host (0) or inside the library by gpp_objective (default)
        self._native = os.environ.get("GPPLUS_NATIVE_OBJECTIVE", "1") != "0"
        return self._fast is not None

    def fun_fast(self, x: np.ndarray, return_grad=True) -> Union[float, Tuple[float, np.ndarray]]:
        """Same value and gradient as ``fun`` without touching the module tree (restart workers only: the
        model's parameters are NOT updated; ``fit_model_scipy`` loads the best theta at the end)."""
        eng = self.model._get_engine()
        self.model._factor_key = None
        if self._native:
            if getattr(eng, "_layout_owner", None) is not self._fast:
                eng.set_theta_layout(self._fast.layout_spec())
                eng._layout_owner = self._fast
            return eng.objective(x, return_grad)  # one GIL-free call: transforms, device evaluation, priors
        return self._fast.fun(x, lambda hyper, want: eng.mll_grad(hyper, want_grad=want), return_grad)

    def fun(self, x: np.ndarray, return_grad=True) -> Union[float, Tuple[float, np.ndarray]]:
        self._load(x)
        self.model.zero_grad()
        obj = -marginal_log_likelihood(self.model, self.add_prior, self.regularization_parameter)
        if return_grad:
            obj.backward()
            return obj.item(), self.pack_grads()
        return obj.item()


def _sample_from_prior(model) -> np.ndarray:
    out = []
    for _, module, prior, closure, _ in model.named_priors():
        if not closure(module).requires_grad:
            continue
        out.append(prior.expand(closure(module).shape).sample().cpu().numpy().ravel())
    return np.concatenate(out)


def get_bounds(likobj, theta):
    """Box bounds per parameter name (mll_scipy.py:149-183)."""
    lo, hi = [], []

    def push(a, b, count):
        lo.extend([a] * count)
        hi.extend([b] * count)

    for name, values in likobj.unpack_parameters(theta).items():
        k = values.numel()
        for col in getattr(likobj.model, "qual_kernel_columns", []):
            if name == str(col):
                push(-3, 3, k)
        if name == "likelihood.noise_covar.raw_noise" or name.startswith("[") or name.startswith("latent["):
            push(-np.inf, np.inf, k)
        if "raw_lengthscale" in name:
            push(-10.0, 3.0, k)
        elif name.startswith("covar_module"):
            push(-10.0, 3.0, k)
        elif name.startswith("mean"):
            push(-1.5, 1.5, k)
        elif name.startswith("Theta_") or name.startswith("encoder"):
            push(-15, 15, k)
        elif name.startswith("A_matrix"):
            push(-10, 10, k)
    return np.array(lo, dtype=float).reshape(-1), np.array(hi, dtype=float).reshape(-1)


def _fit_model_from_state(likobj, theta0, jac, options, method="trust-constr", constraint=False, bounds=False):
    lo, hi = get_bounds(likobj, theta0)
    if constraint is True:
        raise NotImplementedError("the reference's nonlinear latent constraint reads model.nn_model, which GP_Plus "
                                  "does not define (optim/mll_scipy.py:141-147); it is not part of the engine path")
    box = Bounds(lo, hi) if bounds is True else None
    objective = likobj.fun_fast if getattr(likobj, "_fast", None) is not None else likobj.fun
    try:
        with gptsettings.fast_computations(log_prob=False):
            return minimize(fun=objective, x0=theta0, args=(True) if jac else (False), method=method, jac=jac,
                            bounds=box, constraints=[], options=options)
    except Exception as e:
        if isinstance(e, (NotPSDError, NanError)):
            return e  # unstable starting point: scored as +inf by the caller
        if isinstance(e, ValueError) and "within the support" in str(e):
            # torch path only: a float32 softplus underflowed to 0 and torch.distributions rejected the LogNormal
            # prior's argument.  The reference aborts the whole multi-start fit here; one runaway restart is
            # scored +inf instead.
            return NanError(str(e))
        raise


_METHOD_DEFAULTS = {
    "L-BFGS-B": {"ftol": 1e-6, "gtol": 1e-5, "maxfun": 5000, "maxiter": 2000},
    "trust-constr": {"verbose": 1},
    "BFGS": {"gtol": 1e-07, "norm": np.inf, "eps": 1.4901161193847656e-08, "maxiter": None, "disp": False,
             "return_all": False, "finite_diff_rel_step": None},
    "SLSQP": {"maxiter": 100, "ftol": 1e-06, "iprint": 1, "disp": False, "eps": 1.4901161193847656e-08,
              "finite_diff_rel_step": None},
    "Newton-CG": {"xtol": 1e-05, "eps": 1.4901161193847656e-08, "maxiter": None, "disp": False,
                  "return_all": False},
}


def _workers_per_gpu(n_train: int) -> int:
    env = os.environ.get("GPPLUS_WORKERS_PER_GPU")
    if env:
        return max(1, int(env))
    # small and mid-size problems leave most SMs idle in the latency-bound parts of an evaluation (leaf chain, small
    # GEMMs): keep several restarts in flight per GPU.  Measured evals/s at 1 -> 2 -> 4 workers: n=1536 0.8k -> 1.3k ->
    # 1.8k, n=3072 325 -> 537, n=6144 96 -> 120; at n=500 eight workers sustain 8-9k.
    if n_train <= 1024:
        return 8
    if n_train <= 2048:
        return 4
    if n_train <= 8192:
        return 2
    return 1


def _run_restarts(likobj, theta0_list, work, jac, options, method, constraint, bounds, n_jobs) -> Dict[int, object]:
    """Run restarts claimed from ``work`` (a ``parallel.RestartQueue`` shared by every rank) on this process's
    GPUs; returns {index: result} of the restarts this process ran."""
    devices = parallel.local_devices()
    n_train = int(likobj.model.train_targets.shape[0])
    n_workers = len(devices) * _workers_per_gpu(n_train)
    if n_jobs is not None and n_jobs > 0:
        n_workers = min(n_workers, n_jobs)
    n_workers = max(1, min(n_workers, work.count))
    results: Dict[int, object] = {}
    if work.count == 0:
        return results
    errors: List[BaseException] = []
    lock = threading.Lock()

    def worker(slot: int):
        from ..models.gpregression import set_default_device
        set_default_device(devices[slot % len(devices)])
        local = None
        try:
            while True:
                i = work.claim()
                if i is None:
                    break
                if local is None:
                    local = copy.deepcopy(likobj) if n_workers > 1 else likobj
                res = _fit_model_from_state(local, theta0_list[i], jac, options, method, constraint, bounds)
                with lock:
                    results[i] = res
        except BaseException as e:  # propagate non-numerical failures like the reference does
            with lock:
                errors.append(e)
        finally:
            if local is not None and local is not likobj:
                local.model.release_engine()

    if n_workers == 1:
        worker(0)
    else:
        threads = [threading.Thread(target=worker, args=(s,), daemon=True) for s in range(n_workers)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    if errors:
        raise errors[0]
    return results


def _use_lockstep(likobj, method, jac, constraint) -> bool:
    """Small problems (CUDA-graph replay, N <= 2048) with the native objective layout: one host thread per GPU keeps
    up to 64 L-BFGS-B restarts in flight (optim/_lockstep.py) instead of one host thread per restart."""
    if os.environ.get("GPPLUS_LOCKSTEP", "1") == "0":
        return False
    if method != "L-BFGS-B" or jac is not True or constraint is True:
        return False
    if getattr(likobj, "_fast", None) is None or not likobj._native:
        return False
    return int(likobj.model.train_targets.shape[0]) <= 2048


def _run_restarts_lockstep(likobj, theta0_list, work, options, bounds) -> Dict[int, object]:
    from . import _lockstep
    lo = hi = None
    if bounds is True and len(theta0_list) > 0:
        lo, hi = get_bounds(likobj, theta0_list[0])
    devices = parallel.local_devices()
    # a driver fills its slots greedily from the shared queue: cap them at this GPU's fair share so that the first
    # rank / device to start does not claim every restart
    _, world = parallel.world()
    share = -(-work.count // max(1, world * len(devices)))
    n_train = int(likobj.model.train_targets.shape[0])
    cap = max(1, min(_lockstep.slots_for(n_train), share))
    if len(devices) == 1 or work.count <= 1:
        return _lockstep.run_lockstep(likobj, theta0_list, work, options, lo, hi, devices[0], cap)
    results: Dict[int, object] = {}
    errors: List[BaseException] = []
    lock = threading.Lock()

    def drive(dev):
        try:
            out = _lockstep.run_lockstep(likobj, theta0_list, work, options, lo, hi, dev, cap)
            with lock:
                results.update(out)
        except BaseException as e:
            with lock:
                errors.append(e)

    threads = [threading.Thread(target=drive, args=(d,), daemon=True) for d in devices]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results


def fit_model_scipy(
    model,
    add_prior: bool = True,
    num_restarts: int = 1,
    theta0_list: Optional[List[np.ndarray]] = None,
    jac: bool = True,
    options: Dict = {},
    n_jobs: int = -1,
    method="L-BFGS-B",
    constraint=False,
    bounds=False,
    regularization_parameter: List[int] = [0, 0],
) -> Tuple[List[OptimizeResult], float]:
    if method not in _METHOD_DEFAULTS:
        raise ValueError("Wrong method")
    defaults = dict(_METHOD_DEFAULTS[method])
    for key in options.keys():
        if key not in defaults.keys():
            raise RuntimeError("Unknown option %s!" % key)
        defaults[key] = options[key]

    _engine.load_library()  # fail loudly before any work if the CUDA extension is missing
    likobj = MLLObjective(model, add_prior, regularization_parameter)
    likobj.enable_fast_path()

    if theta0_list is None:
        theta0_list = [likobj.pack_parameters()]
        if num_restarts > -1:
            theta0_list.extend([_sample_from_prior(model) for _ in range(num_restarts + 1)])
            theta0_list.pop(0)
    # every rank must optimise from the same list: rank 0's draws win
    theta0_list = parallel.broadcast_theta_list(theta0_list)

    # restarts are claimed from a cross-rank work queue (the reference: joblib's dynamic dispatch, :287-293)
    work = parallel.RestartQueue(len(theta0_list))
    if _use_lockstep(likobj, method, jac, constraint):
        local = _run_restarts_lockstep(likobj, theta0_list, work, defaults, bounds)
    else:
        local = _run_restarts(likobj, theta0_list, work, jac, defaults, method, constraint, bounds, n_jobs)
    out = parallel.gather_restarts(local, len(theta0_list), len(theta0_list[0]) if theta0_list else 0)

    nlls_opt = [np.inf if isinstance(res, Exception) else res.fun for res in out]
    best_idx = int(np.argmin(nlls_opt))
    try:
        likobj._load(out[best_idx].x)
    except Exception:
        pass
    return out, nlls_opt[best_idx]
