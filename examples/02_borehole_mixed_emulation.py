"""Example 02 (Examples/02_Example_Borehole_Mixed_Emulation.ipynb): two categorical inputs with five levels each
are embedded in a 2-D latent space learned with the kernel hyper-parameters."""
import _path  # noqa: F401
from gpplus_b200.models import GP_Plus
from gpplus_b200.preprocessing import train_test_split_normalizeX
from gpplus_b200.test_functions.analytical import borehole_mixed_variables
from gpplus_b200.utils import set_seed

set_seed(4)
qual_dict = {0: 5, 5: 5}
U, y = borehole_mixed_variables(n=10000, qual_dict=qual_dict, random_state=4)
Utrain, Utest, ytrain, ytest = train_test_split_normalizeX(U, y, test_size=0.99, qual_dict=qual_dict)

model = GP_Plus(Utrain, ytrain, qual_dict=qual_dict)
model.fit(bounds=True)

print("latent positions of the 25 level combinations:\n", model.get_latent_space())
model.evaluation(Utest, ytest)
