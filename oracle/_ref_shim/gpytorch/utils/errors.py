class NanError(RuntimeError):
    pass


class NotPSDError(RuntimeError):
    pass


class CachingError(RuntimeError):
    pass
