"""Per-source noise model (likelihoods_noise/multifidelity.py:26-136 of the reference).

One noise variance per data source: the noise diagonal is d_j = sum_i 1[fidel_j == noise_indices[i]]
* noise_i and K_y = K + diag(d).  Here the classes only hold the raw parameters and the source
bookkeeping; the diagonal is added inside the fused covariance kernel (csrc/cov.cuh,
``prep_targets_kernel``) from an int32 group index per point.
"""
import torch

from .._compat import HomoskedasticNoise, _GaussianLikelihoodBase


class Multifidelity_noise(HomoskedasticNoise):
    def __init__(self, noise_prior=None, noise_constraint=None, batch_shape=torch.Size(), num_noises=1):
        super().__init__(noise_prior, noise_constraint, batch_shape, num_tasks=num_noises)

    def group_index(self, fidel_indices, noise_indices):
        """int32 noise group per point; points whose source is not listed get -1 (no noise added, as the
        reference's sum of masked diagonals does)."""
        if fidel_indices is None or len(fidel_indices) == 0:
            raise ValueError("You need to specify a list of indices for noise such as [1,3]")
        fid = torch.as_tensor(fidel_indices).reshape(-1)
        out = torch.full(fid.shape, -1, dtype=torch.int32)
        for i, src in enumerate(noise_indices):
            out[fid == src] = i
        return out


class Multifidelity_likelihood(_GaussianLikelihoodBase):
    def __init__(self, fidel_indices, noise_indices: list = [1], noise_prior=None, noise_constraint=None,
                 learn_additional_noise=False, batch_shape=torch.Size(), **kwargs):
        noise_covar = Multifidelity_noise(noise_prior=noise_prior, noise_constraint=noise_constraint,
                                          batch_shape=batch_shape, num_noises=len(noise_indices))
        super().__init__(noise_covar=noise_covar)
        self.fidel_indices = fidel_indices
        self.noise_indices = noise_indices

    def group_index(self, fidel_indices=None):
        fid = self.fidel_indices if fidel_indices is None else fidel_indices
        return self.noise_covar.group_index(fid, self.noise_indices)
