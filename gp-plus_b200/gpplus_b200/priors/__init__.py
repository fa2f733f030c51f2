"""Prior distributions used by the GP+ models (all host-side, O(p))."""
from .._compat import LogNormalPrior, NormalPrior, Prior
from . import horseshoe as _horseshoe
from . import mollified_uniform as _mollified

LogHalfHorseshoePrior = _horseshoe.LogHalfHorseshoePrior
MollifiedUniformPrior = _mollified.MollifiedUniformPrior

__all__ = ["LogHalfHorseshoePrior", "MollifiedUniformPrior", "NormalPrior", "LogNormalPrior", "Prior"]
