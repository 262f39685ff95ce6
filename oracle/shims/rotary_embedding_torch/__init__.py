"""ORACLE / TEST INFRASTRUCTURE ONLY -- restatement of the published algorithm of
`rotary-embedding-torch` (>=0.8.4, third-party, not in /root/reference, not installed) for the two
symbols the reference uses: `RotaryEmbedding(dim)` (transfusion.py:1499; called with integer
positions at :2296,2411,2621,3223) and `apply_rotary_emb(freqs, t, freqs_seq_dim=-2)` (:965).
freqs = pos (x) theta^(-2i/dim), each frequency repeated for an interleaved (GPT-J style) pair;
rotate_half maps pairs (x0, x1) -> (-x1, x0).  PARITY UNPINNED upstream (no golden vectors);
this restatement is the de-facto oracle.  Not product code."""
import torch
from torch import nn

class RotaryEmbedding(nn.Module):
    def __init__(self, dim, theta = 10000):
        super().__init__()
        freqs = 1. / (theta ** (torch.arange(0, dim, 2)[:(dim // 2)].float() / dim))
        self.freqs = nn.Parameter(freqs, requires_grad = False)

    def forward(self, t, seq_len = None, offset = 0):
        freqs = self.freqs
        freqs = t.type(freqs.dtype)[..., None] * freqs
        return freqs.repeat_interleave(2, dim = -1)

def rotate_half(x):
    x = x.reshape(*x.shape[:-1], -1, 2)
    x1, x2 = x.unbind(dim = -1)
    return torch.stack((-x2, x1), dim = -1).flatten(-2)

def apply_rotary_emb(freqs, t, start_index = 0, scale = 1., seq_dim = -2, freqs_seq_dim = None):
    dtype = t.dtype
    if freqs_seq_dim is None and (freqs.ndim == 2 or t.ndim == 3):
        freqs_seq_dim = 0
    if t.ndim == 3 or freqs_seq_dim is not None:
        seq_len = t.shape[seq_dim]
        idx = [slice(None)] * freqs.ndim
        idx[freqs_seq_dim] = slice(-seq_len, None)
        freqs = freqs[tuple(idx)]
    rot_dim = freqs.shape[-1]
    end_index = start_index + rot_dim
    t_left, t_mid, t_right = t[..., :start_index], t[..., start_index:end_index], t[..., end_index:]
    t_mid = (t_mid * freqs.cos() * scale) + (rotate_half(t_mid) * freqs.sin() * scale)
    return torch.cat((t_left, t_mid, t_right), dim = -1).type(dtype)
