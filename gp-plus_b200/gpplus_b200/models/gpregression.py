"""``GPR``: exact GP regression model backed by the B200 engine.

Mirrors models/gpregression.py:38-175 of the reference (constructor arguments, y min-max scaling,
noise constraint and priors, ``ScaleKernel`` wrapping, ``predict`` contract).  The reference builds
on gpytorch's ``ExactGP``; here the model only owns the raw hyper-parameters and the static training
data, and every O(N^2)/O(N^3) step -- covariance, Cholesky, alpha / log|K|, gradient, predictive
mean / variance -- is a call into ``libgpplus_b200.so``.
"""
from __future__ import annotations

import contextlib

import math
import threading
from typing import List, Optional, Tuple, Union

import numpy as np
import torch

from .. import _engine, kernels
from .._compat import (ConstantMean, GaussianLikelihood, GreaterThan, Kernel, LogNormalPrior, Module,
                       MultivariateNormal, Positive, ZeroMean, settings)
from ..likelihoods_noise.multifidelity import Multifidelity_likelihood
from ..priors import LogHalfHorseshoePrior, MollifiedUniformPrior
from ..utils.transforms import inv_softplus, softplus

_default_device = threading.local()


def set_default_device(index: int):
    """GPU index new engines of this thread are created on (restart workers pin one GPU each)."""
    _default_device.index = int(index)


def get_default_device() -> int:
    return getattr(_default_device, "index", 0)


class _EngineLogProb(torch.autograd.Function):
    """log N(y_scaled; m, K + noise) evaluated on the GPU; backward hands the fused analytic gradient
    (csrc/cov.cuh, grad_tile_kernel) to autograd so the O(p) chain rule to the raw parameters and the
    prior terms stay ordinary torch code (optim/mll_scipy.py:37-60, :112-127)."""

    @staticmethod
    def forward(ctx, model, w, z, sf2, noise, beta):
        eng = model._get_engine()
        hyper = {"w": w.detach().double().cpu().numpy(),
                 "z": z.detach().double().cpu().numpy() if eng.dz > 0 else None,
                 "sigma_f2": float(sf2.detach()),
                 "noise": noise.detach().double().cpu().numpy().reshape(-1),
                 "beta": beta.detach().double().cpu().numpy().reshape(-1) if eng.n_mean > 0 else None}
        need_grad = any(ctx.needs_input_grad[1:])
        out = eng.mll_grad(hyper, want_grad=need_grad)
        model._last_eval = out
        model._factor_key = None  # the engine's factor now belongs to these hyper-parameters (no prediction cache)
        if need_grad:
            ctx.grads = (
                torch.as_tensor(-out["d_w"], dtype=w.dtype).reshape(w.shape),
                torch.as_tensor(-out["d_z"], dtype=z.dtype).reshape(z.shape) if eng.dz > 0 else torch.zeros_like(z),
                torch.as_tensor(-out["d_sigma_f2"], dtype=sf2.dtype).reshape(sf2.shape),
                torch.as_tensor(-out["d_noise"], dtype=noise.dtype).reshape(noise.shape),
                torch.as_tensor(-out["d_beta"], dtype=beta.dtype).reshape(beta.shape) if eng.n_mean > 0
                else torch.zeros_like(beta),
            )
        return torch.tensor(-out["nll"], dtype=torch.float64)

    @staticmethod
    def backward(ctx, g):
        gw, gz, gs, gn, gb = ctx.grads
        g = g.to(torch.float64)
        return (None, (g * gw).to(gw.dtype), (g * gz).to(gz.dtype), (g * gs).to(gs.dtype), (g * gn).to(gn.dtype),
                (g * gb).to(gb.dtype))


class GPR(Module):
    def __init__(
        self,
        train_x: torch.Tensor,
        train_y: torch.Tensor,
        correlation_kernel,
        noise_indices: List[int],
        fix_noise: bool = False,
        fix_noise_val: float = 1e-5,
        lb_noise: float = 1e-12,
    ) -> None:
        if not torch.is_tensor(train_x):
            raise RuntimeError("'train_x' must be a tensor")
        if not torch.is_tensor(train_y):
            raise RuntimeError("'train_y' must be a tensor")
        if train_x.shape[0] != train_y.shape[0]:
            raise RuntimeError("Inputs and output have different number of observations")
        super().__init__()

        noise_constraint = GreaterThan(lb_noise, transform=torch.exp, inv_transform=torch.log)
        if len(noise_indices) == 0:
            likelihood = GaussianLikelihood(noise_constraint=noise_constraint)
        else:
            likelihood = Multifidelity_likelihood(noise_constraint=noise_constraint, noise_indices=noise_indices,
                                                  fidel_indices=train_x[:, -1])
        y_min = train_y.min()
        y_std = train_y.max() - train_y.min()
        train_y_sc = (train_y - y_min) / y_std

        self.train_inputs = (train_x,)
        self.train_targets = train_y_sc
        self.likelihood = likelihood
        self.register_buffer("y_min", y_min)
        self.register_buffer("y_std", y_std)
        self.register_buffer("y_scaled", train_y_sc)
        self._num_outputs = 1

        self.likelihood.register_prior("noise_prior", LogHalfHorseshoePrior(0.01, lb_noise), "raw_noise")
        if fix_noise:
            self.likelihood.raw_noise.requires_grad_(False)
            self.likelihood.noise_covar.noise = torch.tensor(fix_noise_val)

        if isinstance(correlation_kernel, str):
            try:
                kernel_class = getattr(kernels, correlation_kernel)
                correlation_kernel = kernel_class(
                    ard_num_dims=self.train_inputs[0].size(1),
                    lengthscale_constraint=Positive(transform=torch.exp, inv_transform=torch.log),
                )
                correlation_kernel.register_prior(
                    "lengthscale_prior", MollifiedUniformPrior(math.log(0.1), math.log(10)), "raw_lengthscale")
            except Exception:
                raise RuntimeError("%s not an allowed kernel" % correlation_kernel)
        elif not isinstance(correlation_kernel, Kernel):
            raise RuntimeError("specified correlation kernel is not a `gpytorch.kernels.Kernel` instance")

        self.covar_module = kernels.ScaleKernel(
            base_kernel=correlation_kernel,
            outputscale_constraint=Positive(transform=softplus, inv_transform=inv_softplus),
        )
        self.covar_module.register_prior("outputscale_prior", LogNormalPrior(1e-6, 1.0), "outputscale")

        # engine state (never pickled / deep-copied)
        self._engine = None
        self._engine_device = None
        self._factor_key = None
        self._last_eval = None

    # ------------------------------------------------------------------------------------------
    # engine plumbing
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_engine"] = None
        state["_factor_key"] = None
        state["_last_eval"] = None
        return state

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_engine", "_last_eval"):
                new.__dict__[k] = None
            elif k == "_factor_key":
                new.__dict__[k] = None
            else:
                new.__dict__[k] = copy.deepcopy(v, memo)
        return new

    def _quant_columns(self) -> List[int]:
        return list(range(self.train_inputs[0].shape[-1]))

    def _level_index(self, x: torch.Tensor, training: bool) -> Optional[np.ndarray]:
        return None

    def _latent_table(self) -> Optional[torch.Tensor]:
        return None

    def _latent_passes(self) -> int:
        return 1

    def _check_combos_seen(self, levels):
        return None

    def _quant_kernel(self):
        leaves = [k for k in self.covar_module.leaf_kernels() if k.has_lengthscale]
        if len(leaves) != 1:
            raise RuntimeError("GPR expects exactly one stationary kernel under ScaleKernel")
        return leaves[0]

    def _noise_index(self, x: torch.Tensor) -> Optional[np.ndarray]:
        if isinstance(self.likelihood, Multifidelity_likelihood):
            return self.likelihood.group_index(x[:, -1]).numpy()
        return None

    def _mean_layout(self) -> Tuple[int, List[torch.nn.Parameter]]:
        """(n_mean, constants) of the mean model: ZeroMean -> 0, ConstantMean -> 1."""
        mm = getattr(self, "mean_module", None)
        if mm is None or isinstance(mm, ZeroMean):
            return 0, []
        if isinstance(mm, ConstantMean):
            return 1, [mm.constant]
        raise NotImplementedError("mean module %s has no device path" % type(mm).__name__)

    def _mean_index(self, x: torch.Tensor) -> Optional[np.ndarray]:
        return None

    def _engine_kwargs(self, dev: int) -> dict:
        """Constructor arguments of an engine handle holding this model's training set on GPU ``dev``."""
        x = self.train_inputs[0]
        qk = self._quant_kernel() if len(self._quant_columns()) > 0 else None
        family = qk.family if qk is not None else _engine.KERNEL_EXPSQ
        with torch.no_grad():
            table = self._latent_table()
        n_mean, _ = self._mean_layout()
        n_noise = int(self.likelihood.noise_covar.raw_noise.numel())
        xq = x[:, self._quant_columns()].detach().double().cpu().numpy() if qk is not None else None
        return dict(
            xq=xq, y=self.train_targets.detach().double().cpu().numpy(), kernel=family,
            level_idx=self._level_index(x, True), n_combo=0 if table is None else int(table.shape[-2]),
            dz=0 if table is None else int(table.shape[-1]), noise_idx=self._noise_index(x), n_noise=n_noise,
            mean_idx=self._mean_index(x), n_mean=n_mean, device=dev, n_pass=self._latent_passes())

    def _new_engine(self, dev: int, kwargs: Optional[dict] = None) -> "_engine.Engine":
        """A fresh engine handle holding this model's training set on GPU ``dev`` (not cached: the lock-step
        multi-start driver keeps one handle per in-flight restart and passes the same ``kwargs`` to all of them)."""
        return _engine.Engine(**(kwargs if kwargs is not None else self._engine_kwargs(dev)))

    @contextlib.contextmanager
    def engine_override(self, engine):
        """Route every engine call of this model to ``engine`` inside the block (the closed-form objective is validated
        against the torch path through a stub engine, optim/_fast_objective.self_check); the real engine, its
        factorisation cache and the parameters are untouched."""
        prev = getattr(self, "_engine_stub", None)
        self._engine_stub = engine
        try:
            yield engine
        finally:
            self._engine_stub = prev
            self._factor_key = None

    def _get_engine(self) -> "_engine.Engine":
        if getattr(self, "_engine_stub", None) is not None:
            return self._engine_stub
        dev = get_default_device()
        if self._engine is not None and self._engine_device == dev and self._engine.n_pass == self._latent_passes():
            return self._engine
        if self._engine is not None:
            self._engine.close()
        self._engine = self._new_engine(dev)
        self._engine_device = dev
        self._factor_key = None
        return self._engine

    def release_engine(self):
        if self._engine is not None:
            self._engine.close()
        self._engine = None
        self._factor_key = None

    def _natural(self):
        """Differentiable natural hyper-parameters (w, Z, sigma_f^2, noise, beta) from the raw ones."""
        ref = self.covar_module.raw_outputscale
        if len(self._quant_columns()) > 0:
            w = self._quant_kernel().distance_weights()
        else:
            w = torch.zeros(0, dtype=ref.dtype)
        z = self._latent_table()
        if z is None:
            z = torch.zeros(0, dtype=ref.dtype)
        sf2 = self.covar_module.outputscale
        noise = self.likelihood.noise.reshape(-1)
        n_mean, consts = self._mean_layout()
        beta = torch.cat([c.reshape(-1) for c in consts]) if n_mean > 0 else torch.zeros(0, dtype=ref.dtype)
        return w, z, sf2, noise, beta

    def log_marginal(self) -> torch.Tensor:
        """log p(y_scaled | X, theta): differentiable float64 scalar, NOT divided by n."""
        w, z, sf2, noise, beta = self._natural()
        return _EngineLogProb.apply(self, w, z, sf2, noise, beta)

    def prior_mean_and_variance(self):
        """Prior mean vector and prior variance (diag of sigma_f^2 * correlation = sigma_f^2) at the training inputs
        as differentiable float64 tensors -- what ``model(*model.train_inputs)`` returns in training mode
        (gp_plus.py:386-484); used by the interval-score penalty (mll_scipy.py:57-59)."""
        x = self.train_inputs[0]
        n = x.shape[0]
        n_mean, consts = self._mean_layout()
        if n_mean == 0:
            mean = torch.zeros(n, dtype=torch.float64)
        else:
            beta = torch.cat([c.reshape(-1) for c in consts]).to(torch.float64)
            idx = self._mean_index(x)
            if idx is None:
                mean = beta[0].expand(n)
            else:
                idx = torch.as_tensor(idx, dtype=torch.long)
                mean = torch.where(idx >= 0, beta[idx.clamp_min(0)], torch.zeros(n, dtype=torch.float64))
        var = self.covar_module.outputscale.to(torch.float64).reshape(()).expand(n)
        return mean, var

    def _hyper_numpy(self):
        with torch.no_grad():
            w, z, sf2, noise, beta = self._natural()
        eng = self._get_engine()
        return {"w": w.double().numpy(), "z": z.double().numpy() if eng.dz > 0 else None, "sigma_f2": float(sf2),
                "noise": noise.double().numpy(), "beta": beta.double().numpy() if eng.n_mean > 0 else None}

    def _ensure_factor(self):
        eng = self._get_engine()
        hyper = self._hyper_numpy()
        key = tuple(np.concatenate([np.ravel(v) for v in (hyper["w"], hyper["z"] if hyper["z"] is not None else [],
                                                           [hyper["sigma_f2"]], hyper["noise"],
                                                           hyper["beta"] if hyper["beta"] is not None else [])]).tolist())
        if self._factor_key != key:
            eng.factorize(hyper)
            self._factor_key = key
        return eng

    # ------------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor) -> MultivariateNormal:
        """Prior mean and dense covariance at x (off the hot path; kept for API compatibility)."""
        mm = getattr(self, "mean_module", None)
        mean_x = mm(x) if mm is not None else torch.zeros(x.shape[0], dtype=x.dtype)
        covar_x = self.covar_module.outputscale * self.covar_module.base_kernel(x).evaluate()
        return MultivariateNormal(mean_x, covar_x)

    def predict(self, x: torch.Tensor, return_std: bool = False, include_noise: bool = False
                ) -> Union[torch.Tensor, Tuple[torch.Tensor]]:
        """Posterior mean (and standard deviation) in the original y units (gpregression.py:122-149)."""
        self.eval()
        with settings.fast_computations(log_prob=False):
            if self.train_targets.ndim != 1:
                raise NotImplementedError("batched GPs are not supported by the engine")
            if x.dim() != 2:
                raise ValueError("predict expects a 2-D tensor of inputs")
            eng = self._ensure_factor()
            self.fidel_indices = x[:, -1]
            xc = x.detach().cpu()
            cols = self._quant_columns()
            xq = xc[:, cols].double().numpy() if len(cols) > 0 else np.zeros((xc.shape[0], 0))
            add_noise = bool(return_std and include_noise)
            if add_noise and isinstance(self.likelihood, Multifidelity_likelihood):
                self.likelihood.fidel_indices = x[:, -1]
            levels = self._level_index(xc, False)
            if levels is not None:
                self._check_combos_seen(levels)
            mean_sc, var_sc = eng.predict(
                np.ascontiguousarray(xq), level_idx=levels,
                noise_idx=self._noise_index(xc) if add_noise else None, mean_idx=self._mean_index(xc),
                include_noise=add_noise, min_var=settings.min_variance)
            dtype = self.y_std.dtype if self.y_std.dtype.is_floating_point else torch.float64
            out_mean = self.y_min + self.y_std * torch.from_numpy(mean_sc).to(dtype)
            if return_std:
                out_std = torch.from_numpy(var_sc).to(dtype).sqrt() * self.y_std
                return out_mean, out_std
            return out_mean

    def posterior(self, X, output_indices=None, observation_noise=True, posterior_transform=None, **kwargs):
        """Predictive distribution at X in SCALED units (botorch posterior stand-in: mean + marginal variance)."""
        self.eval()
        X = X.double()
        mean, std = self.predict(X, return_std=True, include_noise=bool(observation_noise))
        mu = (mean - self.y_min) / self.y_std
        var = (std / self.y_std) ** 2
        return MultivariateNormal(mu, torch.diag(var))

    def reset_parameters(self) -> None:
        """Reset parameters by sampling from the priors (gpregression.py:168-174)."""
        for _, module, prior, closure, setting_closure in self.named_priors():
            if not closure(module).requires_grad:
                continue
            setting_closure(module, prior.expand(closure(module).shape).sample())
