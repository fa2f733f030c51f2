class GPInputWarning(UserWarning):
    pass


class NumericalWarning(RuntimeWarning):
    pass


class OldVersionWarning(UserWarning):
    pass
