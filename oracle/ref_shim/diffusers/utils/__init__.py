"""diffusers.utils names used by latte_t2v.py:10 (shim; see ../__init__.py)."""
USE_PEFT_BACKEND = False   # 0.24.0 without peft installed: the LoRACompatible* layer classes are used


class BaseOutput:
    """Base of the model-output dataclasses; the reference only needs attribute access (`.sample`)."""

    def to_tuple(self):
        return tuple(self.__dict__.values())


def deprecate(*args, **kwargs):
    return None
