"""AdaLayerNorm / AdaLayerNormZero (shim placeholders): only reached for norm_type 'ada_norm' / 'ada_norm_zero';
Latte-1 uses 'ada_norm_single' (SURVEY.md App. C.2)."""
import torch.nn as nn


class _Unbuilt(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError(f"diffusers shim: {type(self).__name__} is a placeholder (not used by Latte-1)")


class AdaLayerNorm(_Unbuilt):
    pass


class AdaLayerNormZero(_Unbuilt):
    pass
