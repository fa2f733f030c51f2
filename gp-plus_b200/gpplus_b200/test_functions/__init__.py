from .analytical import borehole, borehole_mixed_variables, sine_1D, wing
from .multi_fidelity import Borehole_MF_BO, multi_fidelity_wing, multi_fidelity_wing_value

__all__ = ["wing", "borehole", "borehole_mixed_variables", "multi_fidelity_wing", "multi_fidelity_wing_value",
           "Borehole_MF_BO", "sine_1D"]
