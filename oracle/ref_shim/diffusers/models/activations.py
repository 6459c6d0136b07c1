"""diffusers.models.activations (shim).  GELU = Linear followed by gelu (tanh form when approximate='tanh'); parameter
name `proj` (state-dict key `ff.net.0.proj.*`).  GEGLU / ApproximateGELU exist for isinstance checks only."""
import torch.nn as nn
import torch.nn.functional as F


class GELU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int, approximate: str = "none"):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out)
        self.approximate = approximate

    def forward(self, hidden_states):
        return F.gelu(self.proj(hidden_states), approximate=self.approximate)


class GEGLU(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("diffusers shim: GEGLU is a placeholder (Latte-1 uses gelu-approximate)")


class ApproximateGELU(nn.Module):
    def __init__(self, *a, **k):
        super().__init__()
        raise NotImplementedError("diffusers shim: ApproximateGELU is a placeholder")
