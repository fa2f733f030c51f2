"""Example 05 (Examples/05_Example_1Dsin.ipynb): 1-D sine with a Matern-3/2 kernel, 16 restarts."""
import _path  # noqa: F401
from gpplus_b200.models import GP_Plus
from gpplus_b200.preprocessing import train_test_split_normalizeX
from gpplus_b200.test_functions.analytical import sine_1D
from gpplus_b200.utils import set_seed

set_seed(1)
X, y = sine_1D(n=10000, random_state=1, frequency=1.0, noise_std=0.0)
Xtrain, Xtest, ytrain, ytest = train_test_split_normalizeX(X, y, test_size=0.99)

model = GP_Plus(Xtrain, ytrain, quant_correlation_class="Matern32Kernel")
model.fit(n_jobs=-1, num_restarts=16)
model.evaluation(Xtest, ytest)
model.score(Xtest, ytest, plot_MSE=False)
model.get_params('Omega')
