from .data_type_check import data_type_check
from .interval_score import interval_score_function
from .set_seed import set_seed
from .transforms import inv_softplus, softplus

__all__ = ["set_seed", "data_type_check", "interval_score_function", "softplus", "inv_softplus"]
