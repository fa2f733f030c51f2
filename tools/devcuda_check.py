import sys
sys.path.insert(0,'.'); sys.path.insert(0,'gp-plus_b200')
import warnings, torch
from gpplus_b200.models import GP_Plus
from gpplus_b200.preprocessing import train_test_split_normalizeX
from gpplus_b200.test_functions.analytical import borehole
from gpplus_b200.utils import set_seed
set_seed(1245)
X, y = borehole(n=4000, random_state=12345)
Xtrain, Xtest, ytrain, ytest = train_test_split_normalizeX(X, y, test_size=0.95)
with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter("always")
    model = GP_Plus(Xtrain, ytrain, device='cuda')
    model.fit(n_jobs=-1, num_restarts=3)
    print("warnings:", [str(x.message)[:80] for x in w])
m, s = model.predict(Xtest.to('cuda') if False else Xtest, return_std=True)
print("pred ok", m.shape, float(((m-ytest)**2).mean().sqrt()/ytest.std()))
model.fit(num_restarts=2, optim_type="adam_torch")
print("adam ok")
model.fit(num_restarts=2, optim_type="continuation")
print("continuation ok")
