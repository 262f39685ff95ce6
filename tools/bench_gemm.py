#!/usr/bin/env python
"""Times the tcgen05 GEMM family alone on the shapes of the b128 train step (M = 131072 tokens, d = 512, FFN inner 1365 -> 1408 / 2816).
TFX_LIB=<path to a libtfx_b200 build> selects a library variant, TFX_GEMM_CLUSTER=1|2 the CTA pairing (A/B experiments)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transfusion_pytorch_b200 import _lib
if os.environ.get('TFX_LIB'):
    _lib.LIB_PATH = os.environ['TFX_LIB']

def main():
    ops = _lib.Ops()
    M, D, Ip = int(os.environ.get('TOKENS', 131072)), 512, 1408
    bf = torch.bfloat16
    r = lambda *s: (torch.randn(*s, device = 'cuda') * 0.05).to(bf)
    u, w1, b1 = r(M, D), r(2 * Ip, D), torch.zeros(2 * Ip, device = 'cuda')
    vg, h = torch.empty(M, 2 * Ip, device = 'cuda', dtype = bf), torch.empty(M, Ip, device = 'cuda', dtype = bf)
    w2 = r(D, Ip); dy = r(M, D); dh = torch.empty(M, Ip, device = 'cuda', dtype = bf); du = torch.empty(M, D, device = 'cuda', dtype = bf)
    gw = torch.zeros(2 * Ip, D, device = 'cuda')
    cases = {
        'geglu  [131072 x 2816 x 512]': (lambda: ops.gemm_geglu(u, D, w1, D, b1, M, 2 * Ip, D, vg, h), 2.0 * M * 2 * Ip * D),
        'store  [131072 x 2816 x 512]': (lambda: ops.gemm_store(u, D, 0, w1, D, 0, M, 2 * Ip, D, None, 0, vg, 2 * Ip, None, None, 1.0, 0, 1), 2.0 * M * 2 * Ip * D),
        'dgrad  [131072 x 1408 x 512]': (lambda: ops.gemm_store(dy, D, 0, w2, Ip, 1, M, Ip, D, None, 0, dh, Ip, None, None, 1.0, 0, 1), 2.0 * M * Ip * D),
        'dgrad  [131072 x 512 x 2816]': (lambda: ops.gemm_store(vg, 2 * Ip, 0, w1, D, 1, M, D, 2 * Ip, None, 0, du, D, None, None, 1.0, 0, 1), 2.0 * M * 2 * Ip * D),
        'wgrad  [2816 x 512 x 131072]': (lambda: ops.gemm_store(vg, 2 * Ip, 1, u, D, 1, 2 * Ip, D, M, gw, D, None, 0, None, None, 1.0, 1, 20), 2.0 * M * 2 * Ip * D),
    }
    big = torch.empty(256 << 20, dtype = torch.uint8, device = 'cuda')
    tag = os.path.basename(os.environ.get('TFX_LIB', 'default')) + ' cl=' + os.environ.get('TFX_GEMM_CLUSTER', 'auto')
    for name, (fn, fl) in cases.items():
        for _ in range(3): fn()
        ts = []
        for _ in range(10):
            big.zero_()                                    # flush L2
            e0, e1 = torch.cuda.Event(enable_timing = True), torch.cuda.Event(enable_timing = True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        med = sorted(ts)[5]
        print(f'{tag:36s} {name}: median {med:8.1f} us  min {min(ts):8.1f} us  {fl / med / 1e6:7.1f} TFLOP/s')
if __name__ == '__main__':
    main()
