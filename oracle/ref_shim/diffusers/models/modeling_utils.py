"""ModelMixin (shim): an nn.Module with a dtype/device view; loading from the hub is not restated."""
import torch.nn as nn


class ModelMixin(nn.Module):
    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device
