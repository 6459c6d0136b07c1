"""diffusers.models (shim): latte_t2v.py:9 imports Transformer2DModel but never uses it."""


class Transformer2DModel:
    def __init__(self, *a, **k):
        raise NotImplementedError("diffusers shim: Transformer2DModel is a placeholder")
