"""LoRA-compatible layers without a LoRA branch (shim): plain Linear / Conv2d that accept the `scale` argument the
reference passes (latte_t2v.py:118-121)."""
import torch.nn as nn


class LoRACompatibleLinear(nn.Linear):
    def forward(self, hidden_states, scale: float = 1.0):
        return super().forward(hidden_states)


class LoRACompatibleConv(nn.Conv2d):
    def forward(self, hidden_states, scale: float = 1.0):
        return super().forward(hidden_states)
