#!/usr/bin/env python
"""Times the tcgen05 attention kernels alone on the config-2 shape (32 x 1024 tokens, 8 heads, two 256-token spans per sample).
TFX_LIB=<path to a libtfx_b200 build> selects a library variant (used for A/B experiments on the kernel)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from transfusion_pytorch_b200 import _lib
if os.environ.get('TFX_LIB'):
    _lib.LIB_PATH = os.environ['TFX_LIB']
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from test_ops_gpu import make_rb

def main():
    ops = _lib.Ops()
    B, H, cap, scale = int(os.environ.get('BATCH', 32)), 8, 50., 0.125
    lens = [1024] * B
    spans = [(b, off, 256) for b in range(B) for off in (206, 668)]
    rb = make_rb(lens, spans)
    M = rb.M
    g = torch.Generator(device = 'cuda').manual_seed(1)
    unit = lambda x: (torch.nn.functional.normalize(x.reshape(M, H, 64), dim = -1) * 8.).reshape(M, H * 64).to(torch.bfloat16)
    q, k = unit(torch.randn(M, H * 64, device = 'cuda', generator = g)), unit(torch.randn(M, H * 64, device = 'cuda', generator = g))
    v = torch.randn(M, H * 64, device = 'cuda', generator = g).to(torch.bfloat16)
    gates = torch.randn(M, H, device = 'cuda', generator = g)
    dev = lambda a: torch.from_numpy(a).cuda()
    kvl = dev(rb.kv_limit)
    t2 = [dev(x) for x in (rb.t2_q0, rb.t2_qend, rb.t2_kv0, rb.t2_kvend)]
    k2 = [dev(x) for x in (rb.k2_kv0, rb.k2_kvend, rb.k2_q0, rb.k2_qend, rb.k2_order)]
    fp = torch.zeros(8, device = 'cuda'); z = torch.zeros(64, device = 'cuda')
    ops.attn_fast_params(z, z, 64, scale, cap, fp)
    o = torch.zeros(M, H * 64, device = 'cuda', dtype = torch.bfloat16); lse = torch.zeros(H, M, device = 'cuda')
    do = torch.randn(M, H * 64, device = 'cuda', generator = g).to(torch.bfloat16)
    dsum = torch.zeros(H, M, device = 'cuda'); dq = torch.zeros(M, H * 64, device = 'cuda'); dk = torch.zeros_like(dq); dv = torch.zeros_like(o)
    p2 = dev(rb.p2)
    def fwd(): ops.attn_fwd_tc(q, k, v, H * 64, H * 64, H * 64, gates, H, kvl, *t2, len(rb.t2_q0), o, H * 64, lse, M, 0, scale, cap, fp)
    def fwd_ts(): ops.attn_fwd_ts(q, k, v, H * 64, H * 64, H * 64, gates, H, kvl, *t2, len(rb.t2_q0), p2, len(rb.p2), o, H * 64, lse, M, 0, scale, cap, fp)
    def bwd_ts(): ops.attn_bwd_ts(q, k, v, do, H * 64, H * 64, H * 64, H * 64, lse, dsum, kvl, *k2, len(rb.k2_kv0), dq, dk, dv, H * 64, M, H, scale, cap, fp)
    def bwd(): ops.attn_bwd_tc(q, k, v, do, H * 64, H * 64, H * 64, H * 64, lse, dsum, kvl, *k2, len(rb.k2_kv0), dq, dk, dv, H * 64, M, H, scale, cap, fp)
    big = torch.empty(256 << 20, dtype = torch.uint8, device = 'cuda')
    pairs = float((rb.kv_limit.astype('int64') - rb.cu[:-1].repeat(rb.seq_lens) + 1).sum())
    flops = dict(fwd = 4.0 * pairs * 64 * H, fwd_ts = 4.0 * pairs * 64 * H, bwd = 10.0 * pairs * 64 * H, bwd_ts = 10.0 * pairs * 64 * H)
    for name, fn in (('fwd', fwd), ('fwd_ts', fwd_ts), ('bwd', bwd), ('bwd_ts', bwd_ts)):
        for _ in range(3): fn()
        ts = []
        for _ in range(10):
            big.zero_()                                    # flush L2
            e0, e1 = torch.cuda.Event(enable_timing = True), torch.cuda.Event(enable_timing = True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        med = sorted(ts)[5]
        print(f'{os.environ.get("TFX_LIB", "default"):24s} B={B} {name:7s}: median {med:8.1f} us  min {min(ts):8.1f} us  {flops[name] / med / 1e6:7.1f} TFLOP/s algorithmic')
if __name__ == '__main__':
    main()
