"""gpytorch.lazy (pre-linear_operator): dense stand-ins.  Every "lazy tensor" of this shim holds a dense torch
tensor (diagonal ones hold the diagonal); only the operations the reference performs on them are provided."""
import warnings

import torch

from . import settings
from .utils.errors import NanError, NotPSDError
from .utils.warnings import NumericalWarning


def psd_safe_cholesky(A, upper=False, out=None, jitter=None, max_tries=None):
    """gpytorch.utils.cholesky.psd_safe_cholesky: cholesky_ex, then diagonal jitter 1e-8 (double) / 1e-6 (float)
    times 10^i for i = 0 .. max_tries-1 (default 3), NanError on NaN input, NotPSDError at the end."""
    L, info = torch.linalg.cholesky_ex(A, out=out)
    if not torch.any(info):
        return L.transpose(-1, -2) if upper else L
    isnan = torch.isnan(A)
    if isnan.any():
        raise NanError(f"cholesky_cpu: {isnan.sum().item()} of {A.numel()} elements of the {A.shape} tensor are NaN.")
    if jitter is None:
        jitter = settings.cholesky_jitter.value(A.dtype)
    if max_tries is None:
        max_tries = settings.cholesky_max_tries.value()
    Aprime = A.clone()
    jitter_prev = 0
    for i in range(max_tries):
        jitter_new = jitter * (10 ** i)
        # add jitter only where needed
        diag_add = ((info > 0) * (jitter_new - jitter_prev)).unsqueeze(-1).expand(*Aprime.shape[:-1])
        Aprime.diagonal(dim1=-1, dim2=-2).add_(diag_add)
        jitter_prev = jitter_new
        warnings.warn(f"A not p.d., added jitter of {jitter_new:.1e} to the diagonal", NumericalWarning)
        L, info = torch.linalg.cholesky_ex(Aprime, out=out)
        if not torch.any(info):
            return L.transpose(-1, -2) if upper else L
    raise NotPSDError(f"Matrix not positive definite after repeatedly adding jitter up to {jitter_new:.1e}.")


class LazyTensor:
    """Base: dense matrix holder with the slice of the LazyTensor interface the reference uses."""

    def __init__(self, tensor):
        self.tensor = tensor

    # -- shape --
    @property
    def shape(self):
        return self.tensor.shape

    def size(self, *a):
        return self.tensor.size(*a)

    def dim(self):
        return self.tensor.dim()

    @property
    def dtype(self):
        return self.tensor.dtype

    @property
    def device(self):
        return self.tensor.device

    @property
    def requires_grad(self):
        return self.tensor.requires_grad

    def evaluate(self):
        return self.tensor

    def to_dense(self):
        return self.tensor

    def diag(self):
        return self.tensor.diagonal(dim1=-2, dim2=-1)

    def to(self, *args, **kwargs):
        return self.__class__(self.tensor.to(*args, **kwargs))

    def double(self):
        return self.to(torch.float64)

    def detach(self):
        return self.__class__(self.tensor.detach())

    def __getitem__(self, idx):
        return NonLazyTensor(self.evaluate()[idx])

    def squeeze(self, dim):
        return NonLazyTensor(self.evaluate().squeeze(dim))

    def add_jitter(self, jitter_val=1e-3):
        n = self.shape[-1]
        return NonLazyTensor(self.evaluate() + jitter_val * torch.eye(n, dtype=self.dtype, device=self.device))

    def add_diag(self, diag):
        return NonLazyTensor(self.evaluate() + torch.diag_embed(diag.expand(self.shape[:-1])))

    def __add__(self, other):
        if isinstance(other, LazyTensor):
            return NonLazyTensor(self.evaluate() + other.evaluate())
        return NonLazyTensor(self.evaluate() + other)

    __radd__ = __add__

    def __mul__(self, other):
        if isinstance(other, LazyTensor):
            return NonLazyTensor(self.evaluate() * other.evaluate())
        return NonLazyTensor(self.evaluate() * other)

    __rmul__ = __mul__

    def mul(self, other):
        return self.__mul__(other)

    def matmul(self, other):
        return self.evaluate() @ (other.evaluate() if isinstance(other, LazyTensor) else other)

    def cholesky(self, upper=False):
        return TriangularLazyTensor(psd_safe_cholesky(self.evaluate(), upper=upper), upper=upper)

    def inv_matmul(self, rhs, left_tensor=None):
        L = psd_safe_cholesky(self.evaluate())
        res = torch.cholesky_solve(rhs if rhs.dim() > 1 else rhs.unsqueeze(-1), L)
        if rhs.dim() == 1:
            res = res.squeeze(-1)
        return res if left_tensor is None else left_tensor @ res

    def logdet(self):
        L = psd_safe_cholesky(self.evaluate())
        return 2.0 * torch.log(torch.diagonal(L, dim1=-2, dim2=-1)).sum(-1)

    def inv_quad_logdet(self, inv_quad_rhs=None, logdet=False, reduce_inv_quad=True):
        L = psd_safe_cholesky(self.evaluate())
        iq = None
        if inv_quad_rhs is not None:
            v = torch.linalg.solve_triangular(L, inv_quad_rhs, upper=False)
            iq = (v * v).sum(-2)
            if reduce_inv_quad:
                iq = iq.sum(-1)
        ld = 2.0 * torch.log(torch.diagonal(L, dim1=-2, dim2=-1)).sum(-1) if logdet else None
        return iq, ld


class NonLazyTensor(LazyTensor):
    pass


class TriangularLazyTensor(LazyTensor):
    def __init__(self, tensor, upper=False):
        super().__init__(tensor)
        self.upper = upper


class DiagLazyTensor(LazyTensor):
    """Diagonal matrix (batch) held by its diagonal ``[..., n]``."""

    def __init__(self, diag):
        self._diag = diag

    @property
    def tensor(self):
        return torch.diag_embed(self._diag)

    def diag(self):
        return self._diag

    @property
    def shape(self):
        return torch.Size([*self._diag.shape, self._diag.shape[-1]])

    def size(self, *a):
        return self.shape[a[0]] if a else self.shape

    def dim(self):
        return self._diag.dim() + 1

    @property
    def dtype(self):
        return self._diag.dtype

    @property
    def device(self):
        return self._diag.device

    def to(self, *args, **kwargs):
        return DiagLazyTensor(self._diag.to(*args, **kwargs))

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        if Ellipsis in idx and all((i is Ellipsis) or isinstance(i, (int, slice)) for i in idx):
            # batch indexing such as covar[i, ...] / covar[:, i, ...]: index the batch dimensions of the diagonal
            lead = tuple(i for i in idx if i is not Ellipsis)
            return DiagLazyTensor(self._diag[lead])
        return NonLazyTensor(self.evaluate()[idx])

    def squeeze(self, dim):
        return DiagLazyTensor(self._diag.squeeze(dim))

    def __add__(self, other):
        if isinstance(other, DiagLazyTensor):
            return DiagLazyTensor(self._diag + other._diag)
        return super().__add__(other)

    __radd__ = __add__

    def __iadd__(self, other):
        return self.__add__(other)

    def __mul__(self, other):
        if isinstance(other, DiagLazyTensor):
            return DiagLazyTensor(self._diag * other._diag)
        return super().__mul__(other)

    __rmul__ = __mul__


class ConstantDiagLazyTensor(DiagLazyTensor):
    """``diag_values`` [..., 1] repeated ``diag_shape`` times on the diagonal."""

    def __init__(self, diag_values, diag_shape):
        self.diag_values = diag_values
        self.diag_shape = diag_shape
        super().__init__(diag_values.expand(*diag_values.shape[:-1], diag_shape))

    def to(self, *args, **kwargs):
        return ConstantDiagLazyTensor(self.diag_values.to(*args, **kwargs), self.diag_shape)


def lazify(obj):
    if torch.is_tensor(obj):
        return NonLazyTensor(obj)
    if isinstance(obj, LazyTensor):
        return obj
    raise TypeError("object of class {} cannot be made into a LazyTensor".format(obj.__class__.__name__))


def delazify(obj):
    if torch.is_tensor(obj):
        return obj
    if isinstance(obj, LazyTensor):
        return obj.evaluate()
    raise TypeError("object of class {} cannot be made into a Tensor".format(obj.__class__.__name__))
