import contextlib


@contextlib.contextmanager
def gpt_posterior_settings():
    yield


class fantasize:
    @classmethod
    def on(cls):
        return False


def validate_input_scaling(*args, **kwargs):
    return None
