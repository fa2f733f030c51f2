"""gpytorch.settings: the flags the reference touches (gpytorch/settings.py semantics)."""
import torch


class _feature_flag:
    _default = False
    _state = None

    @classmethod
    def on(cls):
        return cls._default if cls._state is None else cls._state

    @classmethod
    def off(cls):
        return not cls.on()

    @classmethod
    def _set_state(cls, state):
        cls._state = state

    def __init__(self, state=True):
        self.prev = self.__class__._state
        self.state = state

    def __enter__(self):
        self.__class__._set_state(self.state)

    def __exit__(self, *args):
        self.__class__._set_state(self.prev)
        return False


class _fast_covar_root_decomposition(_feature_flag):
    _default = True


class _fast_log_prob(_feature_flag):
    _default = True


class _fast_solves(_feature_flag):
    _default = True


class fast_computations:
    """With every flag off (or N below ``max_cholesky_size``) gpytorch computes through Cholesky; this shim only
    implements that exact path, so the flags are recorded and otherwise ignored."""
    covar_root_decomposition = _fast_covar_root_decomposition
    log_prob = _fast_log_prob
    solves = _fast_solves

    def __init__(self, covar_root_decomposition=True, log_prob=True, solves=True):
        self._cms = (_fast_covar_root_decomposition(covar_root_decomposition), _fast_log_prob(log_prob),
                     _fast_solves(solves))

    def __enter__(self):
        for cm in self._cms:
            cm.__enter__()

    def __exit__(self, *args):
        for cm in self._cms:
            cm.__exit__()
        return False


class trace_mode(_feature_flag):
    _default = False


class debug(_feature_flag):
    _default = True


class fast_pred_var(_feature_flag):
    _default = False


class fast_pred_samples(_feature_flag):
    _default = False


class detach_test_caches(_feature_flag):
    _default = True


class lazily_evaluate_kernels(_feature_flag):
    _default = True


class prior_mode(_feature_flag):
    _default = False


class _value_context:
    _global_value = None

    @classmethod
    def value(cls, *args):
        return cls._global_value

    @classmethod
    def _set_value(cls, value):
        cls._global_value = value

    def __init__(self, value):
        self._orig_value = self.__class__.value()
        self._instance_value = value

    def __enter__(self):
        self.__class__._set_value(self._instance_value)

    def __exit__(self, *args):
        self.__class__._set_value(self._orig_value)
        return False


class max_cholesky_size(_value_context):
    _global_value = 800


class cholesky_max_tries(_value_context):
    _global_value = 3


class _dtype_value_context:
    _global_float_value = None
    _global_double_value = None
    _global_half_value = None

    @classmethod
    def value(cls, dtype):
        if torch.is_tensor(dtype):
            dtype = dtype.dtype
        if dtype == torch.float:
            return cls._global_float_value
        if dtype == torch.double:
            return cls._global_double_value
        if dtype == torch.half:
            return cls._global_half_value
        raise RuntimeError("Unsupported dtype for {}.".format(cls.__name__))


class cholesky_jitter(_dtype_value_context):
    _global_float_value = 1e-6
    _global_double_value = 1e-8


class min_variance(_dtype_value_context):
    _global_float_value = 1e-6
    _global_double_value = 1e-10
    _global_half_value = 1e-3


class min_fixed_noise(_dtype_value_context):
    _global_float_value = 1e-4
    _global_double_value = 1e-6
    _global_half_value = 1e-3
