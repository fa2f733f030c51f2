"""Example 04 (Examples/04_Example_MFBO_Borehole.ipynb): cost-aware multi-fidelity Bayesian optimisation of the
five-source borehole problem.  The notebook passes an undefined name ``Borehole_MF`` as the data generator; the
generator it imports, ``Borehole_MF_BO``, is used here.  ``max_cost`` is reduced so that the script ends quickly."""
import _path  # noqa: F401
import numpy as np

from gpplus_b200.bayesian_optimizations import BO
from gpplus_b200.preprocessing.normalizeX import standard
from gpplus_b200.test_functions.multi_fidelity import Borehole_MF_BO
from gpplus_b200.utils import set_seed

set_seed(0)
qual_dict = {8: 5}
l_bound = [100, 990, 700, 100, 0.05, 10, 1000, 6000]
u_bound = [1000, 1110, 820, 10000, 0.15, 500, 2000, 12000]
n_train_init = {"0": 5, "1": 5, "2": 50, "3": 5, "4": 50}
costs = {"0": 1000, "1": 100, "2": 10, "3": 100, "4": 10}

U_init, y_init = Borehole_MF_BO(True, n_train_init)
U_init, umean, ustd = standard(U_init, qual_dict)
start_cost = sum(costs[str(int(s))] for s in np.asarray(U_init)[:, -1])
bestf, cost = BO(U_init, y_init, costs, l_bound, u_bound, umean.numpy(), ustd.numpy(), qual_dict, Borehole_MF_BO,
                 max_cost=start_cost + 300, fit_kwargs={"num_restarts": 8})
print("incumbent per iteration:", bestf)
print("cumulative cost:", cost)
