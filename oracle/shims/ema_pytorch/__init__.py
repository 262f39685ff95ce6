"""ORACLE / TEST INFRASTRUCTURE ONLY -- minimal stand-in for `ema-pytorch`'s EMA wrapper
(transfusion.py:1687-1697, 2967-2969): deep-copied `ema_model`, `update()` lerp, forwarding of
named methods.  Off the hot path.  Not product code."""
from copy import deepcopy
import torch
from torch import nn

class EMA(nn.Module):
    def __init__(self, model, beta = 0.9999, forward_method_names = (), **kwargs):
        super().__init__()
        self.beta = beta
        self.online_model = [model]
        self.ema_model = deepcopy(model)
        self.ema_model.requires_grad_(False)
        for name in forward_method_names:
            setattr(self, name, getattr(self.ema_model, name))

    @property
    def model(self):
        return self.online_model[0]

    @torch.no_grad()
    def update(self):
        for p_ema, p in zip(self.ema_model.parameters(), self.model.parameters()):
            p_ema.lerp_(p.data, 1. - self.beta)
        for b_ema, b in zip(self.ema_model.buffers(), self.model.buffers()):
            b_ema.copy_(b)

    def forward(self, *args, **kwargs):
        return self.ema_model(*args, **kwargs)
