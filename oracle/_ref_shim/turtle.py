"""Stub: the reference imports names from turtle by accident (likelihoods_noise/multifidelity.py:16,
visual/plot_latent.py:6); the real module needs tkinter."""


def __getattr__(name):
    def _stub(*args, **kwargs):
        raise RuntimeError("turtle.%s stub" % name)
    return _stub
