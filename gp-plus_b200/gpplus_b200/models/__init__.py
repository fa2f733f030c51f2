"""Model classes of the drop-in API: ``GPR`` (exact GP regression) and ``GP_Plus`` (mixed-variable / multi-fidelity)."""
from . import gp_plus as _gp_plus
from . import gpregression as _gpregression

GP_Plus = _gp_plus.GP_Plus
GPR = _gpregression.GPR

__all__ = ["GPR", "GP_Plus"]
