from .horseshoe import LogHalfHorseshoePrior
from .mollified_uniform import MollifiedUniformPrior
from .._compat import LogNormalPrior, NormalPrior, Prior

__all__ = ["LogHalfHorseshoePrior", "MollifiedUniformPrior", "NormalPrior", "LogNormalPrior", "Prior"]
