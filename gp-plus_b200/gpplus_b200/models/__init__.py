from .gpregression import GPR
from .gp_plus import GP_Plus

__all__ = ["GPR", "GP_Plus"]
