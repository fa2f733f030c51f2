from . import errors, warnings  # noqa: F401
