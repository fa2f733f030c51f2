"""Lean host side of ``MLLObjective.fun`` for the restart workers.

``MLLObjective.fun`` (optim/mll_scipy.py:112-127 of the reference) spends its O(p) host work -- loading theta
into the module tree, raw -> natural transforms, log-priors, autograd chain rule, gradient packing -- in ~30
small torch ops, about a millisecond per call.  At N=16384 that is noise next to the 150 ms device evaluation;
at the reference's example sizes (N=100..1000) a device evaluation takes a few hundred microseconds and a
64-restart fit makes ~30 000 calls from GIL-sharing worker threads, so the torch path would bound the fit time.

``FastObjective`` is the same function written out in closed form with numpy: it is *compiled from the model*
(parameter order = ``named_parameters`` with ``requires_grad``, constraints, priors in ``named_priors`` order)
and then validated against the torch path of ``MLLObjective`` on random theta with a stub engine
(``self_check``); any model feature it does not recognise makes ``build`` return ``None`` and the caller keeps
the torch path.  Arithmetic contract: theta is rounded to float32 exactly like the reference (mll_scipy.py:97),
everything else is float64.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Tuple

import numpy as np
import torch

from .._compat import LogNormalPrior, NormalPrior
from ..priors import LogHalfHorseshoePrior, MollifiedUniformPrior

_LN10 = math.log(10.0)
_HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


class _Unsupported(Exception):
    pass


def _softplus(x: float) -> Tuple[float, float]:
    """torch.nn.Softplus(beta=1, threshold=20) and its derivative."""
    if x > 20.0:
        return x, 1.0
    e = math.exp(x)
    return math.log1p(e), e / (1.0 + e)


class FastObjective:
    """``fun(theta) -> (neg log posterior, gradient)`` with the device part delegated to ``engine_call``."""

    def __init__(self, model, add_prior: bool):
        self.model = model
        self.add_prior = bool(add_prior)
        params = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
        self.p = int(sum(max(1, p.numel()) for _, p in params))
        self._slices = {}
        off = 0
        for n, p in params:
            k = max(1, p.numel())
            self._slices[id(p)] = (off, off + k, n)
            off += k
        self._compile_structure()
        self._compile_priors()

    # -- structure -----------------------------------------------------------------------------
    def _slice_of(self, tensor) -> Optional[Tuple[int, int]]:
        ent = self._slices.get(id(tensor))
        return None if ent is None else (ent[0], ent[1])

    def _compile_structure(self):
        m = self.model
        claimed = set()

        def claim(t):
            s = self._slice_of(t)
            if s is not None:
                claimed.add(id(t))
            return s

        # latent map (linear only)
        self.lat = None
        table = m._latent_table()
        if table is not None:
            fm = m._latent_map
            if fm.hidden_num != 0 or fm.fci.bias is not None:
                raise _Unsupported("non-linear latent map")
            w = fm.fci.weight
            self.lat_const = w.detach().double().numpy().copy()
            self.lat = claim(w)
            self.zeta = m.zeta[-1].detach().double().numpy().copy()
            self.lat_ls = float(m._latent_kernel.lengthscale.reshape(-1)[0])
            if m._latent_kernel.raw_lengthscale.requires_grad:
                raise _Unsupported("trainable latent lengthscale")
            self.dz, self.n_onehot = int(w.shape[0]), int(w.shape[1])
        # noise
        lik = m.likelihood
        raw_noise = lik.noise_covar.raw_noise
        cons = lik.noise_covar.raw_noise_constraint
        if getattr(cons, "_transform", None) is not torch.exp:
            raise _Unsupported("noise transform")
        self.noise_lb = float(cons.lower_bound)
        self.noise = claim(raw_noise)
        self.noise_const = raw_noise.detach().double().numpy().reshape(-1).copy()
        # outputscale
        raw_os = m.covar_module.raw_outputscale
        from ..utils.transforms import softplus as _sp
        if getattr(m.covar_module.raw_outputscale_constraint, "_transform", None) is not _sp:
            raise _Unsupported("outputscale transform")
        self.os = claim(raw_os)
        self.os_const = float(raw_os.detach())
        # quantitative kernel
        self.ls = None
        self.dq = len(m._quant_columns())
        if self.dq > 0:
            qk = m._quant_kernel()
            raw_ls = qk.raw_lengthscale
            tr = getattr(qk.raw_lengthscale_constraint, "_transform", None)
            from ..models.gp_plus import _rough_ls
            if tr is torch.exp:
                self.ls_kind = "exp"
            elif tr is _rough_ls:
                self.ls_kind = "rough"
            else:
                raise _Unsupported("lengthscale transform")
            probe = qk.distance_weights().detach().double().numpy().reshape(-1)
            ls = qk.lengthscale.detach().double().numpy().reshape(-1)
            ratio = probe * ls * ls
            if np.allclose(ratio, 0.5):
                self.w_num = 0.5
            elif np.allclose(ratio, 1.0):
                self.w_num = 1.0
            elif np.allclose(probe, ls):
                self.w_num = None  # Rough_RBF class: w = lengthscale
            else:
                raise _Unsupported("distance weights")
            if raw_ls.numel() != self.dq:
                raise _Unsupported("non-ARD lengthscale")
            self.ls = claim(raw_ls)
            self.ls_const = raw_ls.detach().double().numpy().reshape(-1).copy()
        # means
        n_mean, consts = m._mean_layout()
        self.n_mean = n_mean
        self.mean = []
        for c in consts:
            if c.numel() != 1:
                raise _Unsupported("mean constant shape")
            self.mean.append((claim(c), float(c.detach().reshape(-1)[0])))
        # every trainable parameter must have a role
        for pid, (_, _, name) in self._slices.items():
            if pid not in claimed:
                raise _Unsupported("parameter without a device role: %s" % name)
        if getattr(m, "interval_score", False) is True:
            raise _Unsupported("interval score objective")

    def _compile_priors(self):
        """(kind, slice, constants) per prior, in ``named_priors`` order."""
        self.priors: List[Tuple] = []
        if not self.add_prior:
            return
        m = self.model
        for name, module, prior, closure, _ in m.named_priors():
            target = closure(module)
            sl = self._slice_of(target)
            if isinstance(prior, LogNormalPrior):
                if target.data_ptr() != m.covar_module.outputscale.data_ptr() and sl is not None:
                    raise _Unsupported("LogNormal prior on a raw parameter")
                # the only constrained-value prior of GP+: outputscale (gpregression.py:113-115)
                if name.split(".")[-1] != "outputscale_prior":
                    raise _Unsupported("LogNormal prior target")
                self.priors.append(("lognormal_os", None, (float(prior.loc), float(prior.scale))))
                continue
            if sl is None:
                if target.requires_grad:
                    raise _Unsupported("prior on an unknown tensor: %s" % name)
                # frozen parameter: constant contribution
                self.priors.append(("const", None, (float(prior.log_prob(target).sum()),)))
                continue
            if isinstance(prior, NormalPrior):
                loc = np.broadcast_to(prior.loc.detach().double().numpy(), target.shape).reshape(-1).copy()
                sc = np.broadcast_to(prior.scale.detach().double().numpy(), target.shape).reshape(-1).copy()
                self.priors.append(("normal", sl, (loc, sc)))
            elif isinstance(prior, LogHalfHorseshoePrior):
                sc = np.broadcast_to(prior.scale.detach().double().numpy(), target.shape).reshape(-1).copy()
                lb = np.broadcast_to(prior.lb.detach().double().numpy(), target.shape).reshape(-1).copy()
                self.priors.append(("horseshoe", sl, (sc, lb)))
            elif isinstance(prior, MollifiedUniformPrior):
                a = np.broadcast_to(prior.a.detach().double().numpy(), target.shape).reshape(-1).copy()
                b = np.broadcast_to(prior.b.detach().double().numpy(), target.shape).reshape(-1).copy()
                ts = np.broadcast_to(prior.tail_sigma.detach().double().numpy(), target.shape).reshape(-1).copy()
                self.priors.append(("mollified", sl, (a, b, ts)))
            else:
                raise _Unsupported("prior type %s" % type(prior).__name__)

    def layout_spec(self):
        """The same structure as plain data for ``gpp_set_theta_layout`` (include/gpplus_b200.h)."""
        from .. import _engine as E
        spec = {"p": self.p, "off_noise": -1 if self.noise is None else self.noise[0], "noise_const": self.noise_const,
                "noise_lb": self.noise_lb, "off_os": -1 if self.os is None else self.os[0], "os_const": self.os_const}
        if hasattr(self, "zeta"):
            spec.update(off_latent=-1 if self.lat is None else self.lat[0], n_onehot=self.n_onehot, zeta=self.zeta,
                        latent_const=self.lat_const, latent_ls=self.lat_ls)
        if self.dq > 0:
            spec.update(off_ls=-1 if self.ls is None else self.ls[0], ls_const=self.ls_const,
                        ls_kind=1 if self.ls_kind == "rough" else 0, w_num=0.0 if self.w_num is None else self.w_num)
        if self.n_mean > 0:
            spec.update(off_mean=[-1 if s is None else s[0] for s, _ in self.mean],
                        mean_const=[c for _, c in self.mean])
        pri = []
        for kind, sl, c in self.priors:
            if kind == "normal":
                pri.append((E.PRIOR_NORMAL, sl[0], sl[1] - sl[0], c[0], c[1], None))
            elif kind == "lognormal_os":
                pri.append((E.PRIOR_LOGNORMAL_OS, -1, 1, [c[0]], [c[1]], None))
            elif kind == "horseshoe":
                pri.append((E.PRIOR_HORSESHOE, sl[0], sl[1] - sl[0], c[0], c[1], None))
            elif kind == "mollified":
                pri.append((E.PRIOR_MOLLIFIED, sl[0], sl[1] - sl[0], c[0], c[1], c[2]))
            else:
                pri.append((E.PRIOR_CONST, -1, 1, [c[0]], None, None))
        spec["priors"] = pri
        return spec

    # -- evaluation ----------------------------------------------------------------------------
    def natural(self, theta: np.ndarray):
        """theta (already float32-rounded, float64 storage) -> natural hyper-parameters + chain-rule factors."""
        d = {}
        if self.lat is not None or getattr(self, "zeta", None) is not None:
            if hasattr(self, "zeta"):
                A = theta[self.lat[0]:self.lat[1]].reshape(self.dz, self.n_onehot) if self.lat is not None \
                    else self.lat_const
                d["z"] = (self.zeta @ A.T) / self.lat_ls
        raw_n = theta[self.noise[0]:self.noise[1]] if self.noise is not None else self.noise_const
        en = np.exp(raw_n)
        d["noise"] = self.noise_lb + en
        d["_dnoise"] = en
        raw_os = float(theta[self.os[0]]) if self.os is not None else self.os_const
        d["sigma_f2"], d["_dos"] = _softplus(raw_os)
        if self.dq > 0:
            raw_ls = theta[self.ls[0]:self.ls[1]] if self.ls is not None else self.ls_const
            if self.ls_kind == "rough":
                ls = 2.0 ** (-0.5) * np.power(10.0, -raw_ls / 2.0)
                dfac = _LN10          # d w / d raw = ln10 * w  (w ~ ls^-2 ~ 10^raw);  w = ls: -ln10/2 * w
            else:
                ls = np.exp(raw_ls)
                dfac = -2.0           # w ~ ls^-2 = e^(-2 raw);  w = ls: +1 * w
            if self.w_num is None:
                w = ls
                dw = (-0.5 * _LN10 if self.ls_kind == "rough" else 1.0) * w
            else:
                w = self.w_num / (ls * ls)
                dw = dfac * w
            d["w"], d["_dw"] = w, dw
        else:
            d["w"] = np.zeros(0)
        if self.n_mean > 0:
            d["beta"] = np.array([theta[s[0]] if s is not None else c for s, c in self.mean])
        return d

    def _prior_terms(self, theta: np.ndarray, nat, grad: np.ndarray) -> float:
        """sum of log-priors; subtracts their gradient from ``grad`` (objective = nll - sum log p)."""
        total = 0.0
        for kind, sl, c in self.priors:
            if kind == "normal":
                v = theta[sl[0]:sl[1]]
                zed = (v - c[0]) / c[1]
                total += float(np.sum(-0.5 * zed * zed - np.log(c[1]) - _HALF_LOG_2PI))
                grad[sl[0]:sl[1]] -= -zed / c[1]
            elif kind == "lognormal_os":
                s, ds = nat["sigma_f2"], nat["_dos"]
                loc, sc = c
                ls_ = math.log(s)
                zed = (ls_ - loc) / sc
                total += -ls_ - math.log(sc) - _HALF_LOG_2PI - 0.5 * zed * zed
                if self.os is not None:
                    grad[self.os[0]] -= (-1.0 / s - zed / (sc * s)) * ds
            elif kind == "horseshoe":
                v = theta[sl[0]:sl[1]]
                sc, lb = c
                ev = np.exp(v)
                t = lb + ev
                r = sc / t
                u = 1.0 + 3.0 * r * r
                lu = np.log(u)
                total += float(np.sum(np.log(lu) + v))
                grad[sl[0]:sl[1]] -= (6.0 * r / (u * lu)) * (-r * ev / t) + 1.0
            elif kind == "mollified":
                v = theta[sl[0]:sl[1]]
                a, b, ts = c
                mid, half = 0.5 * (a + b), 0.5 * (b - a)
                dev = v - mid
                out = np.maximum(np.abs(dev) - half, 0.0)
                total += float(np.sum(-0.5 * (out / ts) ** 2 - np.log(ts) - _HALF_LOG_2PI
                                      - np.log(1.0 + (b - a) / (math.sqrt(2.0 * math.pi) * ts))))
                grad[sl[0]:sl[1]] -= -(out / (ts * ts)) * np.sign(dev)
            else:  # const
                total += c[0]
        return total

    def fun(self, x: np.ndarray, engine_call: Callable, return_grad: bool = True):
        theta = np.asarray(x, dtype=np.float64).astype(np.float32).astype(np.float64)
        nat = self.natural(theta)
        hyper = {"w": nat["w"], "z": nat.get("z"), "sigma_f2": nat["sigma_f2"], "noise": nat["noise"],
                 "beta": nat.get("beta")}
        out = engine_call(hyper, return_grad)
        grad = np.zeros(self.p)
        if return_grad:
            if self.lat is not None:
                grad[self.lat[0]:self.lat[1]] = ((np.asarray(out["d_z"]).T @ self.zeta) / self.lat_ls).reshape(-1)
            if self.noise is not None:
                grad[self.noise[0]:self.noise[1]] = np.asarray(out["d_noise"]).reshape(-1) * nat["_dnoise"]
            if self.os is not None:
                grad[self.os[0]] = out["d_sigma_f2"] * nat["_dos"]
            if self.ls is not None:
                grad[self.ls[0]:self.ls[1]] = np.asarray(out["d_w"]).reshape(-1) * nat["_dw"]
            for k, (s, _) in enumerate(self.mean):
                if s is not None:
                    grad[s[0]] = np.asarray(out["d_beta"]).reshape(-1)[k]
        logp = self._prior_terms(theta, nat, grad) if self.add_prior else 0.0
        val = float(out["nll"]) - logp
        if return_grad:
            return val, grad
        return val


def build(model, add_prior: bool, regularization_parameter) -> Optional[FastObjective]:
    """FastObjective for ``model`` or ``None`` when the model uses anything outside the closed forms."""
    try:
        if any(float(r) != 0.0 for r in regularization_parameter):
            raise _Unsupported("weight regularisation")
        if not hasattr(model, "_latent_table") or not hasattr(model, "_mean_layout"):
            raise _Unsupported("not an engine-backed model")
        if getattr(model, "embedding_type", "deterministic") != "deterministic":
            raise _Unsupported("probabilistic embedding (sampled latent tables): torch path")
        return FastObjective(model, add_prior)
    except _Unsupported:
        return None


class _StubEngine:
    """Deterministic pseudo-engine: smooth functions of the natural parameters with known gradients, so the
    chain rule and the prior terms of both host paths can be compared without a GPU."""

    def __init__(self, dq, dz, n_combo, n_noise, n_mean, seed=0):
        rng = np.random.RandomState(seed)
        self.dq, self.dz, self.n_combo, self.n_noise, self.n_mean = dq, dz, n_combo, n_noise, n_mean
        self.cw, self.cz = rng.randn(dq), rng.randn(n_combo, max(dz, 1))[:, :dz]
        self.cn, self.cb, self.cs = rng.randn(n_noise), rng.randn(n_mean), rng.randn()

    def mll_grad(self, hyper, want_grad=True):
        w = np.asarray(hyper["w"]).reshape(-1)
        val = float(np.sum(self.cw * np.log1p(w)) + self.cs * math.log(hyper["sigma_f2"])
                    + np.sum(self.cn * np.sqrt(np.asarray(hyper["noise"]).reshape(-1))))
        out = {"d_w": self.cw / (1.0 + w), "d_sigma_f2": self.cs / hyper["sigma_f2"],
               "d_noise": 0.5 * self.cn / np.sqrt(np.asarray(hyper["noise"]).reshape(-1))}
        if self.dz > 0:
            z = np.asarray(hyper["z"]).reshape(self.n_combo, self.dz)
            val += float(np.sum(self.cz * np.sin(z)))
            out["d_z"] = self.cz * np.cos(z)
        if self.n_mean > 0:
            b = np.asarray(hyper["beta"]).reshape(-1)
            val += float(np.sum(self.cb * b * b))
            out["d_beta"] = 2.0 * self.cb * b
        out.update({"nll": val, "logdet": 0.0, "quad": 0.0, "jitter": 0.0})
        return out


def self_check(likobj, fast: FastObjective, trials: int = 2, tol: Optional[float] = None) -> bool:
    """Compare ``fast.fun`` with the torch path ``likobj.fun`` through a stub engine on random theta.

    The closed forms are evaluated in float64.  For a model whose parameters are float32 (the reference's default
    ``dtype=torch.float``) the torch path rounds every transform and prior term to float32 -- e.g. ``lb + exp(raw)``
    loses the 1e-8 noise floor and the horseshoe prior's gradient moves by ~1e-4 relative -- so there the check
    only guards the structure (1e-3); float64 models are held to 1e-9."""
    model = likobj.model
    if tol is None:
        is32 = any(p.dtype == torch.float32 for p in model.parameters())
        tol = 1e-3 if is32 else 1e-9
    table = model._latent_table()
    n_mean, _ = model._mean_layout()
    stub = _StubEngine(len(model._quant_columns()), 0 if table is None else int(table.shape[1]),
                       0 if table is None else int(table.shape[0]),
                       int(model.likelihood.noise_covar.raw_noise.numel()), n_mean)
    # one verdict per model structure: the layout depends on which parameters exist / are frozen, not on their values
    sig = (tuple((n, tuple(p_.shape), str(p_.dtype), bool(p_.requires_grad)) for n, p_ in model.named_parameters()),
           bool(likobj.add_prior), tuple(np.ravel(likobj.regularization_parameter).tolist()), float(tol), int(trials))
    cache = model.__dict__.setdefault("_fast_check_cache", {})
    if sig in cache:
        return cache[sig]
    keep = likobj.pack_parameters().copy()
    rng = np.random.RandomState(1)
    ok = True
    try:
        with model.engine_override(stub):
            for _ in range(trials):
                x = keep + 0.5 * rng.randn(keep.shape[0])
                f_ref, g_ref = likobj.fun(x)
                f, g = fast.fun(x, stub.mll_grad)
                scale = max(1.0, abs(f_ref))
                gscale = max(1.0, float(np.max(np.abs(g_ref))))
                if not (abs(f - f_ref) <= tol * scale and float(np.max(np.abs(g - g_ref))) <= tol * gscale):
                    ok = False
                    break
    except Exception:
        ok = False
    finally:
        likobj._load(keep)
        model._factor_key = None
    cache[sig] = ok
    return ok
