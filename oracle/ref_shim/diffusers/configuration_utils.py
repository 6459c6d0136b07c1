"""ConfigMixin / register_to_config (shim): the decorated __init__'s arguments become `self.config.<name>`
BEFORE the body runs -- latte_t2v.py:571 reads `self.config.sample_size` inside __init__."""
import functools
import inspect
from types import SimpleNamespace


class ConfigMixin:
    @property
    def config(self):
        return self.__dict__["_shim_config"]


def register_to_config(init):
    @functools.wraps(init)
    def wrapper(self, *args, **kwargs):
        sig = inspect.signature(init)
        bound = sig.bind(self, *args, **kwargs)
        bound.apply_defaults()
        values = {k: v for k, v in bound.arguments.items() if k != "self"}
        self.__dict__["_shim_config"] = SimpleNamespace(**values)
        init(self, *args, **kwargs)

    return wrapper
