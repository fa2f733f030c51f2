from .mll_scipy import fit_model_scipy
from .mll_torch import fit_model_torch
from .mll_noise_continuation import fit_model_continuation

__all__ = ["fit_model_scipy", "fit_model_torch", "fit_model_continuation"]
