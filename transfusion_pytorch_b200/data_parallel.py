"""Pure data parallelism for the training step (BASELINE.json: "NCCL allreduce over NVLink on the gradient
buckets only").  One process per GPU; samples are independent units, so the only exchange is the gradient
all-reduce.  The reference has no distributed code of its own - its examples delegate to `accelerate` (DDP),
train_text_only.py:105-128 - this is the B200-native equivalent on the engine's flat gradient buffer.

Buckets: the flat buffer is laid out in `named_parameters()` order (layer 0 first).  Backward finishes layers
from the last to the first, so after layer i's kernels are enqueued the slice holding layers >= i is final and is
all-reduced on a side stream while the remaining layers run (`Engine.backward(bucket_cb=...)`).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class DataParallelTrainer:
    def __init__(self, model, lr = 1e-3, betas = (0.9, 0.999), eps = 1e-8, weight_decay = 0., decoupled_weight_decay = False, overlap = True):
        self.model = model
        self.hp = dict(lr = lr, betas = betas, eps = eps, weight_decay = weight_decay, decoupled = decoupled_weight_decay)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.overlap = overlap and self.world > 1
        self.comm_stream = None
        self._cpu_opt = None

    # ---- engine-backed (CUDA) path
    def _bucket_bounds(self, eng):
        """flat-buffer offset where each layer's parameters start (layers are contiguous in named_parameters order)"""
        if getattr(self, '_bounds', None) is None:
            starts = {}
            for name, off in eng.offs.items():
                if name.startswith('transformer.layers.'):
                    i = int(name.split('.')[2])
                    starts[i] = min(starts.get(i, 1 << 62), off)
            self._bounds = starts
            self._tail = max((off + eng.named[n].numel() for n, off in eng.offs.items() if n.startswith('transformer.layers.')), default = 0)
        return self._bounds

    def step(self, batch, times = None, **fw):
        model = self.model
        eng = model.engine
        cuda = hasattr(eng, 'gflat') or model.device.type == 'cuda'
        if cuda:
            eng.ensure_attached()
            eng.zero_grad()
        else:
            for p in model.parameters():
                p.grad = None
        loss = model(batch, times = times, **fw)
        if cuda and self.overlap:
            bounds = self._bucket_bounds(eng)
            if self.comm_stream is None:
                self.comm_stream = torch.cuda.Stream()
            hi = [self._tail]
            def cb(i):
                lo = bounds[i]
                ev = torch.cuda.Event(); ev.record()
                with torch.cuda.stream(self.comm_stream):
                    self.comm_stream.wait_event(ev)
                    dist.all_reduce(eng.gflat[lo:hi[0]])
                hi[0] = lo
            eng._bucket_cb = cb
            loss.backward()
            eng._bucket_cb = None
            ev = torch.cuda.Event(); ev.record()
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                if hi[0] > 0:
                    dist.all_reduce(eng.gflat[:hi[0]])
                if self._tail < eng.gflat.numel():
                    dist.all_reduce(eng.gflat[self._tail:])
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        else:
            loss.backward()
            if self.world > 1:
                if cuda:
                    dist.all_reduce(eng.gflat)
                else:
                    grads = [p.grad for p in model.parameters() if p.grad is not None]
                    flat = torch._utils._flatten_dense_tensors(grads)
                    dist.all_reduce(flat)
                    for g, f in zip(grads, torch._utils._unflatten_dense_tensors(flat, grads)):
                        g.copy_(f)
        if cuda:
            eng.adam_step(grad_scale = 1.0 / self.world, zero_grads = True, **self.hp)     # next step's zero_grad() is free
        else:
            if self._cpu_opt is None:
                cls = torch.optim.AdamW if self.hp['decoupled'] else torch.optim.Adam
                self._cpu_opt = cls(model.parameters(), lr = self.hp['lr'], betas = self.hp['betas'], eps = self.hp['eps'], weight_decay = self.hp['weight_decay'])
            if self.world > 1:
                for p in model.parameters():
                    if p.grad is not None:
                        p.grad.div_(self.world)
            self._cpu_opt.step()
        return loss
