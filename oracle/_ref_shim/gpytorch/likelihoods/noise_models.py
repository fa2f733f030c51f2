"""gpytorch.likelihoods.noise_models (<= 1.9): homoskedastic noise as a constant diagonal lazy tensor."""
import torch
from torch.nn import Parameter

from ..constraints import GreaterThan
from ..lazy import ConstantDiagLazyTensor, DiagLazyTensor
from ..module import Module


class Noise(Module):
    pass


class _HomoskedasticNoiseBase(Noise):
    def __init__(self, noise_prior=None, noise_constraint=None, batch_shape=torch.Size(), num_tasks=1):
        super().__init__()
        if noise_constraint is None:
            noise_constraint = GreaterThan(1e-4)

        self.register_parameter(name="raw_noise", parameter=Parameter(torch.zeros(*batch_shape, num_tasks)))
        if noise_prior is not None:
            self.register_prior("noise_prior", noise_prior, lambda m: m.noise, lambda m, v: m._set_noise(v))

        self.register_constraint("raw_noise", noise_constraint)

    @property
    def noise(self):
        return self.raw_noise_constraint.transform(self.raw_noise)

    @noise.setter
    def noise(self, value):
        self._set_noise(value)

    def _set_noise(self, value):
        if not torch.is_tensor(value):
            value = torch.as_tensor(value).to(self.raw_noise)
        self.initialize(raw_noise=self.raw_noise_constraint.inverse_transform(value))

    def forward(self, *params, shape=None, **kwargs):
        """In the homoskedastic case the parameters are only used to infer the required shape.  With ``num_tasks``
        noises the result is a batch of ``num_tasks`` constant diagonals, ``[..., 1, num_tasks, n, n]`` -- the layout
        Multifidelity_noise.forward expects (it checks ``covar.shape[1] == len(noise_indices)`` on a 4-D result,
        squeezes the leading 1 and indexes ``covar[i, ...]``; likelihoods_noise/multifidelity.py:82-136)."""
        if shape is None:
            p = params[0] if torch.is_tensor(params[0]) else params[0][0]
            shape = p.shape if len(p.shape) == 1 else p.shape[:-1]
        noise = self.noise
        *batch_shape, n = shape
        noise_batch_shape = noise.shape[:-1] if noise.dim() > 1 else torch.Size()
        num_tasks = noise.shape[-1]
        batch_shape = torch.broadcast_shapes(noise_batch_shape, torch.Size(batch_shape))
        noise = noise.unsqueeze(-2)
        noise_diag = noise.expand(*batch_shape, 1, num_tasks).contiguous()
        if num_tasks == 1:
            noise_diag = noise_diag.view(*batch_shape, 1)
            return ConstantDiagLazyTensor(noise_diag, diag_shape=n)
        return ConstantDiagLazyTensor(noise_diag.unsqueeze(-1), diag_shape=n)


class HomoskedasticNoise(_HomoskedasticNoiseBase):
    def __init__(self, noise_prior=None, noise_constraint=None, batch_shape=torch.Size()):
        super().__init__(noise_prior=noise_prior, noise_constraint=noise_constraint, batch_shape=batch_shape,
                         num_tasks=1)


__all__ = ["Noise", "_HomoskedasticNoiseBase", "HomoskedasticNoise", "DiagLazyTensor"]
