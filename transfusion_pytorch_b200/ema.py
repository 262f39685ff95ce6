"""`Transfusion.create_ema()` (reference transfusion.py:1681-1699: `ema_pytorch.EMA(self, beta, forward_method_names = (...))`).

The EMA copy is a second `Transfusion` whose parameters live in its own flat fp32 buffer with the SAME layout as the online model's,
so `update()` is ONE launch of `tfx_ema_update` over the whole buffer (the per-parameter `lerp_` loop of ema_pytorch).  The contract kept
from `ema_pytorch.EMA`: `.ema_model`, `.model`, `.update()`, `forward(...)` -> `ema_model(...)`, and the forwarded sampling methods
(`sample`, `sample_one`, `sample_many`, `generate_text_only`, `generate_modality_only`).  `update_after_step` / `update_every` default to
the plain every-step lerp that `oracle/shims/ema_pytorch` restates (the package itself is not in the image: parity is against the shim).
"""
from __future__ import annotations

import copy

import torch
from torch import nn


class EMA(nn.Module):
    def __init__(self, model, beta = 0.99, forward_method_names = (), update_after_step = 0, update_every = 1):
        super().__init__()
        self.beta, self.update_after_step, self.update_every = float(beta), int(update_after_step), int(update_every)
        self.online_model = [model]                       # not registered as a sub-module (as in ema_pytorch)
        # the engine (ctypes handles, workspaces, captured graphs) is rebuilt lazily by the copy; only parameters / buffers are copied
        saved = {k: model.__dict__.pop(k) for k in ('_engine', '_last_batch', '_meta_id_cache') if k in model.__dict__}
        model.__dict__['_engine'] = None
        try:
            self.ema_model = copy.deepcopy(model)
        finally:
            model.__dict__.update(saved)
        self.ema_model.requires_grad_(False)
        self.ema_model.eval()
        self.register_buffer('step', torch.zeros((), dtype = torch.long), persistent = True)
        self._steps = 0
        for name in forward_method_names:
            setattr(self, name, self._forwarded(name))

    def _forwarded(self, name):
        def call(*args, **kwargs):
            self._engines()                               # the copy's engine attaches with the online model's parameter layout
            return getattr(self.ema_model, name)(*args, **kwargs)
        call.__name__ = name
        return call

    @property
    def model(self):
        return self.online_model[0]

    def _engines(self):
        src, dst = self.model.engine, self.ema_model.engine
        src.ensure_attached()
        if dst.flat is None:
            # requires_grad is False on the copy: attach exactly the parameters the online engine owns, in the same order (same offsets)
            flags = {n: p.requires_grad for n, p in self.model.named_parameters()}
            for n, p in self.ema_model.named_parameters():
                p.requires_grad_(flags[n])
            dst.ensure_attached()
            self.ema_model.requires_grad_(False)
            for p in self.ema_model.parameters():
                p.grad = None
            dst.gflat = None
        assert dst.flat.numel() == src.flat.numel() and dst.offs == src.offs, 'EMA copy and online model must share one flat parameter layout'
        return src, dst

    @torch.no_grad()
    def copy_params_from_model_to_ema(self):
        src, dst = self._engines()
        dst.flat.copy_(src.flat)
        for b_ema, b in zip(self.ema_model.buffers(), self.model.buffers()):
            b_ema.copy_(b)
        dst.mark_dirty()

    @torch.no_grad()
    def update(self):
        self._steps += 1
        self.step += 1
        if self._steps % self.update_every != 0:
            return
        if self._steps <= self.update_after_step:
            return self.copy_params_from_model_to_ema()
        src, dst = self._engines()
        dst.ops.ema_update(dst.flat, src.flat, dst.flat.numel(), self.beta)       # ema = beta * ema + (1 - beta) * online, one launch
        for b_ema, b in zip(self.ema_model.buffers(), self.model.buffers()):
            b_ema.copy_(b)
        dst.mark_dirty()                                                             # bf16 GEMM operand copies are stale now

    def forward(self, *args, **kwargs):
        self._engines()
        return self.ema_model(*args, **kwargs)
