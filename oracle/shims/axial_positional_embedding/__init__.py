"""ORACLE / TEST INFRASTRUCTURE ONLY -- stand-in for `axial-positional-embedding`'s
ContinuousAxialPositionalEmbedding (transfusion.py:1398-1401, 2795; modality_processing.py:1016,
1039): per-axis MLP on the integer coordinate, summed across axes.  Off the hot path
(add_pos_emb=False in every benchmark config).  Not product code."""
import torch
from torch import nn, tensor

class ContinuousAxialPositionalEmbedding(nn.Module):
    def __init__(self, dim, num_axial_dims, mlp_depth = 2, mlp_expansion = 2.):
        super().__init__()
        self.num_axial_dims = num_axial_dims
        hidden = int(dim * mlp_expansion)
        self.mlps = nn.ModuleList([
            nn.Sequential(nn.Linear(1, hidden), nn.SiLU(), nn.Linear(hidden, dim))
            for _ in range(num_axial_dims)
        ])

    @property
    def device(self):
        return next(self.parameters()).device

    def combine_factorized(self, axial_embeds, axial_dims = None, flatten = False):
        if axial_dims is not None:
            axial_embeds = [e[:int(d)] for e, d in zip(axial_embeds, tuple(axial_dims))]
        out = None
        n = len(axial_embeds)
        for i, e in enumerate(axial_embeds):
            shape = [1] * n + [e.shape[-1]]
            shape[i] = e.shape[0]
            e = e.reshape(shape)
            out = e if out is None else out + e
        if flatten:
            out = out.reshape(-1, out.shape[-1])
        return out

    def forward(self, axial_dims, return_factorized = False, flatten = False):
        if torch.is_tensor(axial_dims):
            axial_dims = axial_dims.tolist()
        embeds = []
        for d, mlp in zip(axial_dims, self.mlps):
            seq = torch.arange(int(d), device = self.device, dtype = torch.float)[:, None]
            embeds.append(mlp(seq))
        if return_factorized:
            return embeds
        return self.combine_factorized(embeds, flatten = flatten)
