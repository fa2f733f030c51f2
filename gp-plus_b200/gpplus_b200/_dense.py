"""Dense covariance matrices for the API-compatibility paths (``kernel(x).evaluate()``,
``model.forward(x)``): evaluated by the same fused sm_100a covariance kernel as the hot path, through a
temporary engine handle.  Off the hot path -- fitting and prediction never materialise K in Python.
"""
import numpy as np
import torch

from . import _engine


def _pick(x, active_dims):
    if active_dims is None:
        return x
    return x.index_select(-1, active_dims.to(torch.long))


def dense_kernel(kernel, x1, x2=None):
    """K(x1, x2) for a stationary leaf kernel or ScaleKernel(leaf).  Cross-covariances are cut out of the
    joint matrix of [x1; x2]."""
    from ._compat import ProductKernel, ScaleKernel
    from .models.gpregression import get_default_device
    sf2 = 1.0
    leaf = kernel
    if isinstance(kernel, ScaleKernel):
        sf2 = float(kernel.outputscale.detach())
        leaf = kernel.base_kernel
    if isinstance(leaf, ProductKernel) or not leaf.has_lengthscale:
        raise NotImplementedError("dense evaluation is provided for stationary leaf kernels; product kernels are "
                                  "evaluated through the owning model (GP_Plus.forward / predict)")
    a = _pick(x1.detach().cpu().double(), leaf.active_dims)
    n1 = a.shape[0]
    pts = a if x2 is None else torch.cat([a, _pick(x2.detach().cpu().double(), leaf.active_dims)], 0)
    with torch.no_grad():
        w = leaf.distance_weights().double().reshape(-1)
    if w.numel() == 1 and pts.shape[1] > 1:
        w = w.expand(pts.shape[1])
    eng = _engine.Engine(xq=np.ascontiguousarray(pts.numpy()), y=np.zeros(pts.shape[0]), kernel=leaf.family,
                         n_noise=1, n_mean=0, device=get_default_device())
    try:
        K = eng.covariance({"w": w.numpy(), "sigma_f2": sf2, "noise": np.ones(1)})
    finally:
        eng.close()
    K = torch.from_numpy(K)
    return K if x2 is None else K[:n1, n1:].contiguous()


def dense_model_covariance(model, x):
    """sigma_f^2 * k_latent * k_quant at the rows of x (original column layout of the model)."""
    from .models.gpregression import get_default_device
    x = x.detach().cpu()
    cols = model._quant_columns()
    hyper = None
    with torch.no_grad():
        w, z, sf2, noise, beta = model._natural()
    lvl = model._level_index(x, model.training)
    eng = _engine.Engine(xq=np.ascontiguousarray(x[:, cols].double().numpy()) if len(cols) else None,
                         y=np.zeros(x.shape[0]), kernel=model._quant_kernel().family if len(cols) else 0,
                         level_idx=lvl, n_combo=0 if lvl is None else int(z.shape[0]),
                         dz=0 if lvl is None else int(z.shape[1]), n_noise=1, n_mean=0, device=get_default_device())
    try:
        hyper = {"w": w.double().numpy(), "z": z.double().numpy() if lvl is not None else None,
                 "sigma_f2": float(sf2), "noise": np.ones(1)}
        K = eng.covariance(hyper)
    finally:
        eng.close()
    return torch.from_numpy(K)
