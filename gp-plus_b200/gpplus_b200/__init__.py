"""gpplus_b200 -- B200-native exact-GP engine behind GP+'s Python API.

The package mirrors the reference's module layout for the hot path only (models, kernels,
likelihoods_noise, priors, optim, bayesian_optimizations, utils, preprocessing, test_functions);
all O(N^2)/O(N^3) arithmetic runs in ``lib/libgpplus_b200.so`` (hand-written sm_100a CUDA, C ABI in
``include/gpplus_b200.h``).  There is no CPU fallback.
"""
__version__ = "0.1.0"
