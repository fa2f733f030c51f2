"""Stub of botorch (<= 0.7): the reference imports base classes and a posterior wrapper only
(models/gpregression.py:28-33)."""
from . import settings  # noqa: F401
