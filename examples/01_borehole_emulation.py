"""Example 01 of the reference (Examples/01_Example_Borehole_Emulation.ipynb) on the B200 engine:
8-D borehole function, n = 500 training points, rough-RBF kernel, 32-restart MAP fit."""
import _path  # noqa: F401
from gpplus_b200.models import GP_Plus
from gpplus_b200.preprocessing import train_test_split_normalizeX
from gpplus_b200.test_functions.analytical import borehole
from gpplus_b200.utils import set_seed

set_seed(1245)
X, y = borehole(n=10000, random_state=12345)
Xtrain, Xtest, ytrain, ytest = train_test_split_normalizeX(X, y, test_size=0.95)

model = GP_Plus(Xtrain, ytrain)
model.fit(n_jobs=-1, num_restarts=32)

model.evaluation(Xtest, ytest)
noise = model.likelihood.noise_covar.noise.detach() * model.y_std ** 2
print("noise variance", noise, "outputscale", model.covar_module.outputscale.item())
print("omegas", model.covar_module.base_kernel.raw_lengthscale.detach())
