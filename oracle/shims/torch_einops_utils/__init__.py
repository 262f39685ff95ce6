"""ORACLE / TEST INFRASTRUCTURE ONLY -- restatement of the helpers lucidrains/transfusion-pytorch
imports from `torch-einops-utils` (>=0.1.12, not installed): transfusion.py:48-58,
modality_processing.py:47-51.  Not product code."""
from functools import wraps
import torch
import torch.nn.functional as F
from torch.utils._pytree import tree_map
from einops import pack, unpack

def pack_with_inverse(t, pattern):
    is_one = torch.is_tensor(t)
    ts = [t] if is_one else t
    packed, shapes = pack(ts, pattern)
    def inverse(out, inv_pattern = None):
        outs = unpack(out, shapes, inv_pattern if inv_pattern is not None else pattern)
        return outs[0] if is_one else outs
    return packed, inverse

def pad_at_dim(t, pad, dim = -1, value = 0.):
    dims_from_right = (-dim - 1) if dim < 0 else (t.ndim - dim - 1)
    return F.pad(t, (0, 0) * dims_from_right + tuple(pad), value = value)

def pad_left_at_dim(t, n, dim = -1, value = 0.):
    return pad_at_dim(t, (n, 0), dim = dim, value = value)

def pad_right_at_dim(t, n, dim = -1, value = 0.):
    return pad_at_dim(t, (0, n), dim = dim, value = value)

def pad_sequence(tensors, dim = -1, value = 0., left = False, dim_stack = 0, return_lens = False):
    lens = [t.shape[dim] for t in tensors]
    max_len = max(lens)
    padded = [pad_at_dim(t, (max_len - l, 0) if left else (0, max_len - l), dim = dim, value = value) for t, l in zip(tensors, lens)]
    out = torch.stack(padded, dim = dim_stack)
    if return_lens:
        return out, torch.tensor(lens, device = out.device)
    return out

def batched_index_select(t, indices):
    # t [b, m, ...], indices [b, m'] -> [b, m', ...]
    b = t.shape[0]
    batch = torch.arange(b, device = t.device).reshape(b, *((1,) * (indices.ndim - 1)))
    return t[batch, indices]

def reverse_cumsum(t, dim = -1):
    return t.flip(dims = (dim,)).cumsum(dim = dim).flip(dims = (dim,))

def tree_map_tensor(fn, tree):
    return tree_map(lambda x: fn(x) if torch.is_tensor(x) else x, tree)

def tree_map_tensor_to_device(tree, device):
    return tree_map_tensor(lambda x: x.to(device), tree)

def temp_eval(fn):
    @wraps(fn)
    def inner(self, *args, **kwargs):
        was_training = self.training
        self.eval()
        try:
            return fn(self, *args, **kwargs)
        finally:
            self.train(was_training)
    return inner
