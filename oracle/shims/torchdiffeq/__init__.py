"""ORACLE / TEST INFRASTRUCTURE ONLY -- restatement of `torchdiffeq.odeint` (third-party, unpinned,
not installed) for the one configuration the reference uses: method='midpoint', fixed grid equal
to the requested time points, atol/rtol ignored (transfusion.py:1314-1318, call sites :2039,2525,
2911).  y_{k+1} = y_k + h f(t_k + h/2, y_k + h/2 f(t_k, y_k)).  PARITY UNPINNED upstream.
Not product code."""
import torch

def odeint(func, y0, t, *, rtol = 1e-7, atol = 1e-9, method = None, options = None, **_):
    assert method in ('midpoint', 'euler', None) or True
    ys = [y0]
    y = y0
    for t0, t1 in zip(t[:-1], t[1:]):
        dt = t1 - t0
        if method == 'euler':
            y = y + dt * func(t0, y)
        else:
            f0 = func(t0, y)
            y_mid = y + f0 * (0.5 * dt)
            y = y + dt * func(t0 + 0.5 * dt, y_mid)
        ys.append(y)
    return torch.stack(ys)
