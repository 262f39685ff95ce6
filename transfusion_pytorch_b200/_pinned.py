"""Grow-only pool of pinned host staging buffers for the per-step H2D copies (token metadata, pageable latents).

`cudaHostAlloc` costs milliseconds, so buffers are reused; a buffer is handed out again only after the CUDA event recorded
behind the copies that read it has completed, so callers that never synchronise simply get a fresh buffer.
"""
from __future__ import annotations

import torch


class PinnedPool:
    def __init__(self):
        self._idle = []                     # [(uint8 tensor, event)]

    def take(self, nbytes: int) -> torch.Tensor:
        for i, (buf, ev) in enumerate(self._idle):
            if buf.numel() >= nbytes and ev.query():
                del self._idle[i]
                return buf
        return torch.empty(max(int(nbytes), 256), dtype = torch.uint8).pin_memory()

    def give(self, buf: torch.Tensor) -> None:
        """call after the async copies out of `buf` have been enqueued on the current stream"""
        ev = torch.cuda.Event()
        ev.record()
        self._idle.append((buf, ev))
        if len(self._idle) > 16:
            self._idle = [(b, e) for b, e in self._idle if not e.query()][-8:] + [(b, e) for b, e in self._idle if e.query()][-8:]


POOL = PinnedPool()
