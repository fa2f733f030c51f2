"""Optimisers over the B200 engine; names as in the reference's ``gpplus.optim``."""
from . import mll_noise_continuation as _continuation
from . import mll_scipy as _scipy
from . import mll_torch as _adam

fit_model_scipy = _scipy.fit_model_scipy                        # multi-start MAP fit with scipy optimisers
fit_model_torch = _adam.fit_model_torch                         # Adam on the per-point likelihood
fit_model_continuation = _continuation.fit_model_continuation   # noise-continuation ladder

__all__ = ["fit_model_scipy", "fit_model_torch", "fit_model_continuation"]
