"""gpytorch.functions.RBFCovariance: custom-autograd fast path of the non-ARD RBF kernel.  Same value as the plain
formula ``exp(-0.5 * sq_dist(x1 / l, x2 / l))``; evaluated here with ordinary autograd."""


class RBFCovariance:
    @staticmethod
    def apply(x1, x2, lengthscale, sq_dist_func):
        x1_ = x1.div(lengthscale)
        x2_ = x2.div(lengthscale)
        unitless_sq_dist = sq_dist_func(x1_, x2_)
        return unitless_sq_dist.div(-2.0).exp()
