"""ORACLE / TEST INFRASTRUCTURE ONLY -- restatement of the tiny subset of `einx` (>=0.3.0, not
installed, no network) that lucidrains/transfusion-pytorch calls (transfusion.py:196,200,409-410,
443,446-447,450,463,466-467,633,1011,3184).  Elementwise ops over named-axis patterns:
every operand is aligned to the output axis order, missing axes broadcast.  Not product code."""
import torch

def _tokens(side):
    return [t for t in side.strip().split(' ') if t]

def _expand_ellipsis(names, ndim):
    if '...' not in names:
        return names
    i = names.index('...')
    n_extra = ndim - (len(names) - 1)
    return names[:i] + [f'_e{k}' for k in range(n_extra)] + names[i + 1:]

def _elementwise(pattern, operands, fn):
    if '->' in pattern:
        lhs, rhs = pattern.split('->')
    else:
        lhs, rhs = pattern, None
    in_specs = [_tokens(s) for s in lhs.split(',')]
    assert len(in_specs) == len(operands), (pattern, len(operands))
    ops = []
    for spec, op in zip(in_specs, operands):
        if not torch.is_tensor(op):
            assert len(spec) == 0, f'python scalar operand must have an empty pattern: {pattern}'
            ops.append((spec, op)); continue
        ops.append((_expand_ellipsis(spec, op.ndim), op))
    if rhs is None:
        out_names = max((s for s, _ in ops), key = len)
    else:
        out_names = _tokens(rhs)
        if '...' in out_names:
            longest = max((s for s, _ in ops), key = len)
            extra = [n for n in longest if n.startswith('_e')]
            i = out_names.index('...')
            out_names = out_names[:i] + extra + out_names[i + 1:]
    aligned = []
    for spec, op in ops:
        if not torch.is_tensor(op):
            aligned.append(op); continue
        assert len(spec) == op.ndim, (pattern, spec, op.shape)
        named = [n for n in out_names if n in spec]
        perm = [spec.index(n) for n in named]
        t = op.permute(*perm) if perm else op
        shape, k = [], 0
        for n in out_names:
            if n in spec:
                shape.append(t.shape[k]); k += 1
            else:
                shape.append(1)
        aligned.append(t.reshape(shape))
    return fn(*aligned)

def _mk(fn):
    def op(pattern, *operands):
        return _elementwise(pattern, operands, fn)
    return op

less          = _mk(lambda a, b: a < b)
greater       = _mk(lambda a, b: a > b)
greater_equal = _mk(lambda a, b: a >= b)
less_equal    = _mk(lambda a, b: a <= b)
equal         = _mk(lambda a, b: a == b)
logical_and   = _mk(lambda a, b: a & b)
logical_or    = _mk(lambda a, b: a | b)
multiply      = _mk(lambda a, b: a * b)
add           = _mk(lambda a, b: a + b)
subtract      = _mk(lambda a, b: a - b)

def _where(c, a, b):
    ref = a if torch.is_tensor(a) else b
    if not torch.is_tensor(a): a = torch.as_tensor(a, dtype = ref.dtype, device = ref.device)
    if not torch.is_tensor(b): b = torch.as_tensor(b, dtype = ref.dtype, device = ref.device)
    return torch.where(c, a, b)

where = _mk(_where)
