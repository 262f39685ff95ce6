"""Shared helpers of the parity tests: rebuild the exact inputs the golden fixtures were generated from."""
import os

import torch

from transfusion_pytorch_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, f'{name}.pt'), weights_only = False)


def golden_inputs(name):
    """(batch, times) exactly as oracle/make_golden.py built them."""
    if name == 'small_one_modality':
        return synth.small_batch(3, seed = 1, dim_latent = 32, text_vocab = 64)
    if name == 'small_two_modalities':
        return synth.config4_batch(2, seed = 2, total_len = 300, dims = (32, 16), text_vocab = 64)
    if name == 'config2_b2':
        return synth.config2_batch(2, seed = 4)
    raise KeyError(name)


def golden_noise(fx, batch, dim_latents):
    """Per-type noise tensors in the order the reference's flat strategy drew them (one randn_like per type,
    types in order of first appearance)."""
    order = []
    for s in batch:
        for p in s:
            t = p[0] if isinstance(p, tuple) else (0 if (torch.is_tensor(p) and p.is_floating_point()) else None)
            if t is not None and t not in order:
                order.append(t)
    noise = [None] * len(dim_latents)
    for k, t in enumerate(order):
        rows, dl = fx['noise_shapes'][k]
        assert dl == dim_latents[t]
        noise[t] = torch.randn(rows, dl, generator = torch.Generator().manual_seed(9000 + k + 17 * fx['seed']))
    return noise


def grad_fingerprint(named_grads):
    out = {}
    for name, g in named_grads:
        g = g.detach().float().reshape(-1).cpu()
        proj = torch.randn(g.numel(), generator = torch.Generator().manual_seed(1234))
        out[name] = dict(stats = torch.stack([g.sum(), g.abs().sum(), (g * proj).sum(), g.norm()]).double(), head = g[:8].clone())
    return out


def unpack_rows(packed, rb, width = None):
    """packed [M, d] -> padded [B, n_max, d] like the reference's batch layout"""
    n_max = int(rb.seq_lens.max())
    d = packed.shape[1] if width is None else width
    out = packed.new_zeros((rb.B, n_max, d))
    for b in range(rb.B):
        out[b, :rb.seq_lens[b]] = packed[rb.cu[b]:rb.cu[b + 1], :d]
    return out
