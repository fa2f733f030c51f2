"""Example 03 (Examples/03_Example_Wing_multi_fidelity.ipynb): four fidelity sources of the wing-weight function
fused in one GP: the source index is a categorical input, every source has its own noise and mean constant."""
import _path  # noqa: F401
from gpplus_b200.models import GP_Plus
from gpplus_b200.preprocessing import train_test_split_normalizeX
from gpplus_b200.test_functions.multi_fidelity import multi_fidelity_wing
from gpplus_b200.utils import set_seed

set_seed(4)
qual_dict = {10: 4}
num = {"0": 5000, "1": 10000, "2": 10000, "3": 10000}
noise_std = {"0": 0.5, "1": 1.0, "2": 1.5, "3": 2.0}
X, y = multi_fidelity_wing(n=num, noise_std=noise_std, random_state=4)
Xtrain, Xtest, ytrain, ytest = train_test_split_normalizeX(X, y, test_size=0.99, qual_dict=qual_dict,
                                                           stratify=X[..., list(qual_dict.keys())])

model = GP_Plus(Xtrain, ytrain, qual_dict=qual_dict, multiple_noise=True, m_gp="multiple_constant")
model.fit(n_jobs=-1)

model.evaluation(Xtest, ytest)
print("latent positions of the four sources:\n", model.get_latent_space())
print("noise variance per source:", model.noise_value())
