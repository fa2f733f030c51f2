"""Stub of matplotlib.pyplot: every function is a no-op."""


def __getattr__(name):
    def _noop(*args, **kwargs):
        return None
    return _noop
