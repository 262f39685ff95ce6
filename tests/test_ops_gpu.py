"""GPU: individual C-ABI kernels against a plain PyTorch fp32 reference of the same op."""
import math

import numpy as np
import pytest
import torch

from transfusion_pytorch_b200 import _lib
from transfusion_pytorch_b200.modality_processing import RaggedBatch, build_tiles

pytestmark = pytest.mark.gpu
BF16, F32 = torch.bfloat16, torch.float32


@pytest.fixture(scope = 'module')
def ops():
    return _lib.Ops()


@pytest.fixture(params = [1, 2], ids = ['single', 'paired'])
def cluster_mode(ops, request):
    """GEMM launches as independent CTAs (default) and as 2-CTA clusters sharing the B tile by TMA multicast"""
    assert ops.lib.tfx_gemm_set_cluster_mode(request.param) == 0
    yield request.param
    ops.lib.tfx_gemm_set_cluster_mode(3)


@pytest.mark.parametrize('a_mn,b_mn', [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize('M,N,K', [(300, 390, 520), (1000, 1664, 512), (128, 128, 64), (40000, 512, 192)])      # the last one: several tiles per CTA (pair), odd tile count
def test_gemm_store_all_majors(ops, cluster_mode, a_mn, b_mn, M, N, K):
    g = torch.Generator(device = 'cuda').manual_seed(0)
    A = torch.randn(M, K, device = 'cuda', generator = g).to(BF16)
    B = torch.randn(N, K, device = 'cuda', generator = g).to(BF16)
    ref = A.float() @ B.float().t()
    pad8 = lambda x: (x + 7) // 8 * 8
    def store(mat, mn):                      # mat is [MN, K]; returns (tensor, ld)
        if mn:
            t = torch.zeros(K, pad8(mat.shape[0]), device = 'cuda', dtype = BF16); t[:, :mat.shape[0]] = mat.t(); return t, t.shape[1]
        t = torch.zeros(mat.shape[0], pad8(K), device = 'cuda', dtype = BF16); t[:, :K] = mat; return t, t.shape[1]
    (a, lda), (b, ldb) = store(A, a_mn), store(B, b_mn)
    out = torch.zeros(M, N, device = 'cuda', dtype = F32)
    outb = torch.zeros(M, pad8(N), device = 'cuda', dtype = BF16)
    bias = torch.randn(N, device = 'cuda')
    ops.gemm_store(a, lda, a_mn, b, ldb, b_mn, M, N, K, out, N, outb, pad8(N), bias, None, 0.5, 0, 1)
    torch.cuda.synchronize()
    want = 0.5 * ref + bias
    assert torch.allclose(out, want, atol = 2e-2, rtol = 1e-3)
    assert torch.allclose(outb[:, :N].float(), want, atol = 0.15, rtol = 2e-2)
    acc = torch.ones(M, N, device = 'cuda', dtype = F32)
    ops.gemm_store(a, lda, a_mn, b, ldb, b_mn, M, N, K, acc, N, None, 0, None, None, 1.0, 1, 4)
    torch.cuda.synchronize()
    assert torch.allclose(acc, ref + 1, atol = 2e-2, rtol = 1e-3)


def dense_attention(q, k, v, gates, kv_limit, cu, scale, cap):
    out = torch.zeros_like(q, dtype = F32)
    H = q.shape[1] // 64
    for b in range(len(cu) - 1):
        s, e = cu[b], cu[b + 1]
        qq, kk, vv = (t[s:e].float().reshape(e - s, H, 64).transpose(0, 1) for t in (q, k, v))
        sim = torch.einsum('hid,hjd->hij', qq * scale, kk)
        sim = torch.tanh(sim / cap) * cap
        j = torch.arange(s, e, device = q.device)
        mask = j[None, :] <= kv_limit[s:e, None]
        sim = sim.masked_fill(~mask[None], -1e30)
        o = torch.einsum('hij,hjd->hid', sim.softmax(-1), vv)
        o = o * torch.sigmoid(gates[s:e].t())[..., None]
        out[s:e] = o.transpose(0, 1).reshape(e - s, H * 64)
    return out


def make_rb(lens, spans):
    cu = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    M = int(cu[-1])
    kv = np.arange(M, dtype = np.int32); qf = np.arange(M, dtype = np.int32)
    for b, off, ln in spans:
        kv[cu[b] + off: cu[b] + off + ln] = cu[b] + off + ln - 1
        qf[cu[b] + off: cu[b] + off + ln] = cu[b] + off
    z = np.zeros(M, dtype = np.int32)
    rb = RaggedBatch(B = len(lens), M = M, seq_lens = np.asarray(lens, dtype = np.int64), cu = cu, full_lens = np.asarray(lens), text_id = z, label = z, kv_limit = kv,
                     rope_pos = z, cond_row = z, slot = z, n_cond = 0, cond_times = np.zeros(0, np.float32), n_types = 0, type_rows = [], row_token = np.zeros(0, np.int32),
                     row_time = np.zeros(0, np.float32), latents = [], instances = [], modality_positions = [], total_tokens = M, n_type_tokens = [])
    build_tiles(rb, qf)
    return rb


@pytest.mark.parametrize('lens,spans', [([1024, 1024], [(0, 206, 256), (0, 668, 256), (1, 100, 700)]), ([77, 130, 5], [(0, 10, 40), (1, 64, 64), (1, 128, 2)]), ([64], [])])
def test_attention_forward_backward_vs_dense(ops, lens, spans):
    H, cap, scale = 4, 50., 0.125
    rb = make_rb(lens, spans)
    M = rb.M
    g = torch.Generator(device = 'cuda').manual_seed(1)
    q, k, v = (torch.randn(M, H * 64, device = 'cuda', generator = g).to(BF16) * 2 for _ in range(3))
    gates = torch.randn(M, H, device = 'cuda', generator = g)
    dev = lambda a: torch.from_numpy(a).cuda()
    kvl = dev(rb.kv_limit)
    o = torch.zeros(M, H * 64, device = 'cuda', dtype = BF16); lse = torch.zeros(H, M, device = 'cuda')
    ops.attn_fwd(q, k, v, H * 64, H * 64, H * 64, gates, H, kvl, dev(rb.tile_q0), dev(rb.tile_qend), dev(rb.tile_kv0), dev(rb.tile_kvend), len(rb.tile_q0), o, H * 64, lse, M, scale, cap, None)
    qf, kf, vf, gf = (t.float().requires_grad_(True) for t in (q, k, v, gates))
    ref = dense_attention(qf, kf, vf, gf, kvl.long(), rb.cu.tolist(), scale, cap)
    torch.cuda.synchronize()
    assert torch.allclose(o.float(), ref, atol = 3e-2, rtol = 3e-2)
    do = torch.randn(M, H * 64, device = 'cuda', generator = g).to(BF16)
    ref.backward(do.float())
    dop = torch.zeros_like(do); dsum = torch.zeros(H, M, device = 'cuda'); dsum2 = torch.zeros(M, H, device = 'cuda')
    ops.attn_bwd_prep(do, o, gates, dop, dsum, dsum2, None, M, H)
    dq = torch.zeros(M, H * 64, device = 'cuda'); dk = torch.zeros(M, H * 64, device = 'cuda'); dv = torch.zeros(M, H * 64, device = 'cuda', dtype = BF16)
    ops.attn_bwd(q, k, v, dop, H * 64, H * 64, H * 64, H * 64, lse, dsum, kvl, dev(rb.kt_kv0), dev(rb.kt_kvend), dev(rb.kt_q0), dev(rb.kt_qend), len(rb.kt_kv0), dq, dk, dv, H * 64,
                 M, H, scale, cap, None)
    torch.cuda.synchronize()
    for ours, want, name in ((dq, qf.grad, 'dq'), (dk, kf.grad, 'dk'), (dv.float(), vf.grad, 'dv')):
        err = (ours - want).abs().max().item() / want.abs().max().item()
        assert err < 4e-2, (name, err)
    dgate_ref = gf.grad
    dgate = (1 - torch.sigmoid(gates)) * dsum2
    assert (dgate - dgate_ref).abs().max().item() / dgate_ref.abs().max().item() < 4e-2


@pytest.mark.parametrize('lens,spans', [([1024, 1024], [(0, 206, 256), (0, 668, 256), (1, 100, 700)]), ([77, 130, 5, 300], [(0, 10, 40), (1, 64, 64), (1, 128, 2), (3, 120, 150)]), ([64], []),
                                        ([128, 129, 127], [(1, 0, 129)]), ([640] * 6 + [385, 1000, 257], [(b, 100, 300) for b in range(6)] + [(7, 100, 800)])])
def test_attention_tcgen05_fast_path_vs_dense(ops, lens, spans):
    """bounded-logit tcgen05 forward (attention_sm100.cu): RMS-normalised q/k as the QKVG epilogue produces them"""
    H, cap, scale = 4, 50., 0.125
    rb = make_rb(lens, spans)
    M = rb.M
    g = torch.Generator(device = 'cuda').manual_seed(1)
    def unit(x):
        x = x.reshape(M, H, 64)
        return (torch.nn.functional.normalize(x, dim = -1) * 8.).reshape(M, H * 64).to(BF16)
    q, k = (unit(torch.randn(M, H * 64, device = 'cuda', generator = g)) for _ in range(2))
    q[: M // 2] = k[: M // 2]                       # aligned q/k: logits reach the bound (|s| = 8)
    v = (torch.randn(M, H * 64, device = 'cuda', generator = g) * 2).to(BF16)
    gates = torch.randn(M, H, device = 'cuda', generator = g)
    dev = lambda a: torch.from_numpy(a).cuda()
    kvl = dev(rb.kv_limit)
    fp = torch.zeros(8, device = 'cuda')
    zeros = torch.zeros(64, device = 'cuda')
    ops.attn_fast_params(zeros, zeros, 64, scale, cap, fp)
    torch.cuda.synchronize()
    assert fp[0].item() == 1.0 and 8.0 <= fp[1].item() <= 8.2
    o = torch.zeros(M, H * 64, device = 'cuda', dtype = BF16); lse = torch.zeros(H, M, device = 'cuda')
    ops.attn_fwd_tc(q, k, v, H * 64, H * 64, H * 64, gates, H, kvl, dev(rb.t2_q0), dev(rb.t2_qend), dev(rb.t2_kv0), dev(rb.t2_kvend), len(rb.t2_q0), o, H * 64, lse, M, 0, scale, cap, fp)
    ref = dense_attention(q.float(), k.float(), v.float(), gates, kvl.long(), rb.cu.tolist(), scale, cap)
    # the general kernel must skip when the fast flag is set, and agree when run
    o2 = torch.zeros_like(o); lse2 = torch.zeros_like(lse)
    ops.attn_fwd(q, k, v, H * 64, H * 64, H * 64, gates, H, kvl, dev(rb.tile_q0), dev(rb.tile_qend), dev(rb.tile_kv0), dev(rb.tile_kvend), len(rb.tile_q0), o2, H * 64, lse2, M, scale, cap, fp)
    torch.cuda.synchronize()
    assert (o2 == 0).all()
    ops.attn_fwd(q, k, v, H * 64, H * 64, H * 64, gates, H, kvl, dev(rb.tile_q0), dev(rb.tile_qend), dev(rb.tile_kv0), dev(rb.tile_kvend), len(rb.tile_q0), o2, H * 64, lse2, M, scale, cap, None)
    torch.cuda.synchronize()
    assert torch.allclose(o.float(), ref, atol = 3e-2, rtol = 3e-2)
    assert torch.allclose(lse, lse2, atol = 2e-3, rtol = 1e-4)
    assert torch.allclose(o.float(), o2.float(), atol = 2e-2, rtol = 2e-2)
    # ---- backward: tcgen05 kernel vs autograd of the dense reference and vs the general kernel
    qf, kf, vf, gf = (t.float().requires_grad_(True) for t in (q, k, v, gates))
    ref = dense_attention(qf, kf, vf, gf, kvl.long(), rb.cu.tolist(), scale, cap)
    do = torch.randn(M, H * 64, device = 'cuda', generator = g).to(BF16)
    ref.backward(do.float())
    dop = torch.zeros_like(do); dsum = torch.zeros(H, M, device = 'cuda'); dsum2 = torch.zeros(M, H, device = 'cuda')
    dq = torch.full((M, H * 64), 7., device = 'cuda')              # cleared by the prep kernel
    ops.attn_bwd_prep(do, o, gates, dop, dsum, dsum2, dq, M, H)
    dk = torch.zeros(M, H * 64, device = 'cuda'); dv = torch.zeros(M, H * 64, device = 'cuda', dtype = BF16)
    ops.attn_bwd_tc(q, k, v, dop, H * 64, H * 64, H * 64, H * 64, lse, dsum, kvl, dev(rb.k2_kv0), dev(rb.k2_kvend), dev(rb.k2_q0), dev(rb.k2_qend), dev(rb.k2_order), len(rb.k2_kv0), dq, dk, dv, H * 64,
                    M, H, scale, cap, fp)
    dq2 = torch.zeros(M, H * 64, device = 'cuda'); dk2 = torch.zeros(M, H * 64, device = 'cuda'); dv2 = torch.zeros(M, H * 64, device = 'cuda', dtype = BF16)
    ops.attn_bwd(q, k, v, dop, H * 64, H * 64, H * 64, H * 64, lse, dsum, kvl, dev(rb.kt_kv0), dev(rb.kt_kvend), dev(rb.kt_q0), dev(rb.kt_qend), len(rb.kt_kv0), dq2, dk2, dv2, H * 64,
                 M, H, scale, cap, None)
    # transposed-score kernel (attention_bwd_sm100.cu): same contract, P^T / dS^T as TMEM operands
    dq3 = torch.zeros(M, H * 64, device = 'cuda'); dk3 = torch.full((M, H * 64), 3., device = 'cuda'); dv3 = torch.zeros(M, H * 64, device = 'cuda', dtype = BF16)
    for _ in range(2):                              # twice: relaunch on dirty outputs (dq accumulates and is cleared by the caller, dk / dv are overwritten)
        dq3.zero_()
        ops.attn_bwd_ts(q, k, v, dop, H * 64, H * 64, H * 64, H * 64, lse, dsum, kvl, dev(rb.k2_kv0), dev(rb.k2_kvend), dev(rb.k2_q0), dev(rb.k2_qend), dev(rb.k2_order), len(rb.k2_kv0), dq3, dk3, dv3,
                        H * 64, M, H, scale, cap, fp)
    torch.cuda.synchronize()
    for ours, gen, want, name in ((dq, dq2, qf.grad, 'dq'), (dk, dk2, kf.grad, 'dk'), (dv.float(), dv2.float(), vf.grad, 'dv'),
                                  (dq3, dq2, qf.grad, 'dq_ts'), (dk3, dk2, kf.grad, 'dk_ts'), (dv3.float(), dv2.float(), vf.grad, 'dv_ts')):
        err = (ours - want).abs().max().item() / want.abs().max().item()
        err2 = (ours - gen).abs().max().item() / want.abs().max().item()
        assert err < 4e-2 and err2 < 4e-2, (name, err, err2)
    # large gammas: the fast path must decline
    ops.attn_fast_params(zeros + 2.0, zeros + 2.0, 64, scale, cap, fp)
    torch.cuda.synchronize()
    assert fp[0].item() == 0.0


@pytest.mark.parametrize('lens,spans', [([1024, 1024], [(0, 206, 256), (0, 668, 256), (1, 100, 700)]), ([77, 130, 5, 300], [(0, 10, 40), (1, 64, 64), (1, 128, 2), (3, 120, 150)]), ([64], []),
                                        ([128, 129, 127], [(1, 0, 129)]), ([1024] * 40, [(b, 206, 256) for b in range(40)] + [(b, 668, 256) for b in range(40)]),
                                        ([385, 1000, 257, 640, 129, 900, 31], [(1, 100, 800), (3, 0, 640), (5, 300, 77), (5, 500, 300)])])
def test_attention_persistent_ts_forward_vs_dense(ops, lens, spans):
    """persistent two-warpgroup forward with P in TMEM (attention_fwd_sm100.cu) vs the dense fp32 attention and vs the round-1 tcgen05 kernel;
    the 40-sequence case gives every CTA several work items (Q double buffering, K / V ring wrap-around, group drift across items)"""
    H, cap, scale = 4, 50., 0.125
    rb = make_rb(lens, spans)
    M = rb.M
    g = torch.Generator(device = 'cuda').manual_seed(1)
    def unit(x):
        x = x.reshape(M, H, 64)
        return (torch.nn.functional.normalize(x, dim = -1) * 8.).reshape(M, H * 64).to(BF16)
    q, k = (unit(torch.randn(M, H * 64, device = 'cuda', generator = g)) for _ in range(2))
    q[: M // 2] = k[: M // 2]
    v = (torch.randn(M, H * 64, device = 'cuda', generator = g) * 2).to(BF16)
    gates = torch.randn(M, H, device = 'cuda', generator = g)
    dev = lambda a: torch.from_numpy(a).cuda()
    kvl = dev(rb.kv_limit)
    fp = torch.zeros(8, device = 'cuda')
    zeros = torch.zeros(64, device = 'cuda')
    ops.attn_fast_params(zeros, zeros, 64, scale, cap, fp)
    t2 = [dev(getattr(rb, n)) for n in ('t2_q0', 't2_qend', 't2_kv0', 't2_kvend')]
    # pair table: every 128-row tile appears exactly once
    seen = sorted([c >> 1 for c in rb.p2.tolist()] + [(c >> 1) + 1 for c in rb.p2.tolist() if c & 1])
    assert seen == list(range(len(rb.t2_q0)))
    o = torch.zeros(M, H * 64, device = 'cuda', dtype = BF16); lse = torch.zeros(H, M, device = 'cuda')
    for _ in range(2):                              # twice: a relaunch must not depend on leftover state
        o.zero_(); lse.zero_()
        ops.attn_fwd_ts(q, k, v, H * 64, H * 64, H * 64, gates, H, kvl, *t2, len(rb.t2_q0), dev(rb.p2), len(rb.p2), o, H * 64, lse, M, 0, scale, cap, fp)
    o1 = torch.zeros_like(o); lse1 = torch.zeros_like(lse)
    ops.attn_fwd_tc(q, k, v, H * 64, H * 64, H * 64, gates, H, kvl, *t2, len(rb.t2_q0), o1, H * 64, lse1, M, 0, scale, cap, fp)
    torch.cuda.synchronize()
    assert torch.allclose(lse, lse1, atol = 1e-4, rtol = 1e-5)
    assert torch.allclose(o.float(), o1.float(), atol = 1e-2, rtol = 1e-2)
    if M <= 4096:
        ref = dense_attention(q.float(), k.float(), v.float(), gates, kvl.long(), rb.cu.tolist(), scale, cap)
        assert torch.allclose(o.float(), ref, atol = 3e-2, rtol = 3e-2)


def test_rowops_vs_torch(ops):
    M, D, nc = 777, 512, 5
    g = torch.Generator(device = 'cuda').manual_seed(2)
    x = torch.randn(M, D, device = 'cuda', generator = g) * 3 + 1
    cond_row = torch.randint(-1, nc, (M,), device = 'cuda', generator = g, dtype = torch.int32)
    film = torch.randn(nc, 2 * D, device = 'cuda', generator = g) * 0.3
    gam = torch.randn(D, device = 'cuda', generator = g) * 0.3
    u = torch.zeros(M, D, device = 'cuda', dtype = BF16); stats = torch.zeros(M, 2, device = 'cuda')
    ops.adaln_fwd(x, cond_row, film, 2 * D, gam, u, stats, M, D)
    xh = torch.nn.functional.layer_norm(x, (D,))
    cr = cond_row.long().clamp(min = 0)
    want = torch.where((cond_row >= 0)[:, None], xh * (film[cr, :D] + 1) + film[cr, D:], xh * (gam + 1))
    torch.cuda.synchronize()
    assert torch.allclose(u.float(), want, atol = 3e-2, rtol = 1e-2)
    # attention residual
    hid = [torch.randn(M, D, device = 'cuda', generator = g) for _ in range(5)]
    pq = torch.randn(D, device = 'cuda', generator = g) * 0.5
    import ctypes
    arr = (ctypes.c_void_p * 5)(*[h.data_ptr() for h in hid])
    xo = torch.zeros(M, D, device = 'cuda')
    ops.attn_residual_fwd(ctypes.cast(arr, ctypes.c_void_p), 5, gam, pq, xo, None, None, M, D)
    vals = torch.stack(hid)
    keys = torch.nn.functional.normalize(vals, dim = -1) * D ** 0.5 * (gam + 1)
    sim = torch.einsum('lnd,d->nl', keys, pq) * D ** -0.5
    want = torch.einsum('nl,lnd->nd', sim.softmax(-1), vals)
    torch.cuda.synchronize()
    assert torch.allclose(xo, want, atol = 1e-4, rtol = 1e-4)


def test_ce_and_mse_heads_vs_torch(ops):
    M, V = 500, 390
    g = torch.Generator(device = 'cuda').manual_seed(3)
    logits = torch.randn(M, 392, device = 'cuda', generator = g) * 3
    labels = torch.randint(-1, V, (M,), device = 'cuda', generator = g, dtype = torch.int32)
    dl = torch.zeros(M, 392, device = 'cuda', dtype = BF16); acc = torch.zeros(1, device = 'cuda', dtype = torch.float64); nv = torch.zeros(1, device = 'cuda', dtype = torch.int32)
    ops.ce_fwd_bwd(logits, 392, labels, V, 0, 0.25, dl, 392, acc, nv, M)
    lg = logits[:, :V].clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lg, labels.long(), ignore_index = -1, reduction = 'sum')
    (ref * 0.25).backward()
    torch.cuda.synchronize()
    assert abs(acc.item() - ref.item()) / ref.item() < 1e-5 and nv.item() == int((labels >= 0).sum())
    assert torch.allclose(dl[:, :V].float(), lg.grad, atol = 2e-3, rtol = 1e-2) and (dl[:, V:] == 0).all()


# ================================================================================================ fused GEMM epilogues vs fp32 torch
def _rope_tables(ops, n_pos):
    freqs = 1. / (10000 ** (torch.arange(0, 64, 2, device = 'cuda').float() / 64))
    t = torch.empty(n_pos, 32, 2, device = 'cuda'); tt = torch.empty(32, n_pos, 2, device = 'cuda')
    ops.rope_table(freqs, t, tt, n_pos, 32)
    return freqs, t, tt


def _rope_ref(x, pos, freqs):                       # interleaved pairs (x0, x1) -> (x0 c - x1 s, x1 c + x0 s)
    ang = (pos[:, None].float() * freqs).repeat_interleave(2, dim = -1)            # [M, 64]
    x2 = x.reshape(*x.shape[:-1], 32, 2)
    rot = torch.stack((-x2[..., 1], x2[..., 0]), dim = -1).flatten(-2)
    return x * ang.cos()[:, None] + rot * ang.sin()[:, None]


@pytest.mark.parametrize('M,H,D', [(700, 8, 512), (300, 2, 128), (257, 4, 256)])
def test_gemm_qkvg_epilogue_vs_torch(ops, cluster_mode, M, H, D):
    """to_qk | to_v | to_gates GEMM + per-head qk-RMSNorm + interleaved RoPE (T.py:946-965, 1027), incl. the kv-cache row scatter"""
    g = torch.Generator(device = 'cuda').manual_seed(4)
    HI, NQ = H * 64, 3 * H * 64 + 128
    u = torch.randn(M, D, device = 'cuda', generator = g).to(BF16)
    W = torch.zeros(NQ, D, device = 'cuda', dtype = BF16)
    W[:3 * HI + H] = (torch.randn(3 * HI + H, D, device = 'cuda', generator = g) / D ** 0.5).to(BF16)
    gq, gk = (torch.randn(64, device = 'cuda', generator = g) * 0.3 for _ in range(2))
    pos = torch.randint(0, 900, (M,), device = 'cuda', generator = g, dtype = torch.int32)
    freqs, t, tt = _rope_tables(ops, 1024)
    q, k, v = (torch.zeros(M, HI, device = 'cuda', dtype = BF16) for _ in range(3))
    gates = torch.zeros(M, H, device = 'cuda'); inv = torch.zeros(M, 2 * H, device = 'cuda')
    ops.gemm_qkvg(u, D, W, D, M, H, D, q, k, v, gates, inv, gq, gk, pos, tt, 1024, None, None)
    y = u.float() @ W.float().t()
    rms = lambda x, gm: torch.nn.functional.normalize(x, dim = -1) * 8. * (gm + 1.)
    qr = _rope_ref(rms(y[:, :HI].reshape(M, H, 64), gq), pos, freqs).reshape(M, HI)
    kr = _rope_ref(rms(y[:, HI:2 * HI].reshape(M, H, 64), gk), pos, freqs).reshape(M, HI)
    torch.cuda.synchronize()
    assert torch.allclose(q.float(), qr, atol = 6e-2, rtol = 2e-2) and torch.allclose(k.float(), kr, atol = 6e-2, rtol = 2e-2)
    assert torch.allclose(v.float(), y[:, 2 * HI:3 * HI], atol = 3e-2, rtol = 2e-2)
    assert torch.allclose(gates, y[:, 3 * HI:3 * HI + H], atol = 2e-2, rtol = 1e-2)
    want_inv = 1. / y[:, :2 * HI].reshape(M, 2 * H, 64).norm(dim = -1)
    assert torch.allclose(inv, want_inv, rtol = 1e-2, atol = 1e-4)
    # kv-cache append: k / v rows land at kv_rows[m] of a larger matrix, q stays dense
    rows = torch.randperm(2 * M + 50, device = 'cuda', generator = g)[:M].to(torch.int32)
    kc = torch.zeros(2 * M + 50, HI, device = 'cuda', dtype = BF16); vc = torch.zeros_like(kc)
    q2 = torch.zeros_like(q)
    ops.gemm_qkvg(u, D, W, D, M, H, D, q2, kc, vc, gates, inv, gq, gk, pos, tt, 1024, rows, None)
    torch.cuda.synchronize()
    assert torch.equal(q2, q) and torch.equal(kc[rows.long()], k) and torch.equal(vc[rows.long()], v)
    untouched = torch.ones(2 * M + 50, dtype = torch.bool, device = 'cuda'); untouched[rows.long()] = False
    assert (kc[untouched] == 0).all() and (vc[untouched] == 0).all()


@pytest.mark.parametrize('M,N,K,two', [(900, 512, 512, False), (333, 512, 1408, False), (500, 512, 1024, True), (130, 128, 128, False)])
def test_gemm_resid_epilogue_vs_torch(ops, cluster_mode, M, N, K, two):
    """branch output projection + AdaptiveWrapper output gate + residual (T.py:765-769, 1031, 1238-1242); `two`: skip_proj on cat(x, skip) (T.py:1217-1219)"""
    g = torch.Generator(device = 'cuda').manual_seed(5)
    nc = 4
    A = torch.randn(M, K, device = 'cuda', generator = g).to(BF16)
    W = (torch.randn(N, K, device = 'cuda', generator = g) / K ** 0.5).to(BF16)
    bias = torch.randn(N, device = 'cuda', generator = g) * 0.2
    x_res = torch.randn(M, N, device = 'cuda', generator = g)
    x_out = torch.zeros(M, N, device = 'cuda'); xb = torch.zeros(M, N, device = 'cuda', dtype = BF16); yb = torch.zeros(M, N, device = 'cuda', dtype = BF16)
    y = A.float() @ W.float().t()
    if two:
        A1, A2 = A[:, :K // 2].contiguous(), A[:, K // 2:].contiguous()
        ops.gemm_resid(A1, K // 2, A2, K // 2, K // 2, W, K, M, N, K, None, x_res, x_out, xb, None, None, None, 0, None)
        torch.cuda.synchronize()
        want = x_res + y
        assert torch.allclose(x_out, want, atol = 3e-2, rtol = 1e-2) and torch.allclose(xb.float(), want, atol = 6e-2, rtol = 2e-2)
        return
    cond_row = torch.randint(-1, nc, (M,), device = 'cuda', generator = g, dtype = torch.int32)
    zg = torch.rand(nc, 3 * N, device = 'cuda', generator = g)                       # strided table: row pitch 3N, this wrapper's slice starts at column N
    ls = torch.randn(N, device = 'cuda', generator = g) * 0.3
    ops.gemm_resid(A, K, None, 0, 0, W, K, M, N, K, bias, x_res, x_out, None, yb, cond_row, zg[:, N:], 3 * N, ls)
    torch.cuda.synchronize()
    yy = y + bias
    cr = cond_row.long().clamp(min = 0)
    scale = torch.where((cond_row >= 0)[:, None], zg[cr, N:2 * N], ls + 1.)
    assert torch.allclose(yb.float(), yy, atol = 6e-2, rtol = 2e-2)
    assert torch.allclose(x_out, x_res + yy * scale, atol = 4e-2, rtol = 1e-2)
    # text-only form (no condition table): layerscale on every row
    ops.gemm_resid(A, K, None, 0, 0, W, K, M, N, K, bias, x_res, x_out, None, None, None, None, 0, ls)
    torch.cuda.synchronize()
    assert torch.allclose(x_out, x_res + yy * (ls + 1.), atol = 4e-2, rtol = 1e-2)


@pytest.mark.parametrize('M,D,inner', [(600, 512, 1365), (200, 128, 341)])
def test_gemm_geglu_epilogue_and_backward_vs_torch(ops, cluster_mode, M, D, inner):
    """FeedForward net.0 + GEGLU (T.py:833-834, 846-847) on the tile-interleaved W1, and tfx_geglu_bwd against autograd"""
    g = torch.Generator(device = 'cuda').manual_seed(6)
    Ip = (inner + 63) // 64 * 64
    W1 = torch.randn(2 * inner, D, device = 'cuda', generator = g) / D ** 0.5          # rows [value 0..inner) | gate inner..2 inner)  (T.py:833)
    b1 = torch.randn(2 * inner, device = 'cuda', generator = g) * 0.3
    u = torch.randn(M, D, device = 'cuda', generator = g).to(BF16)
    Wp = torch.zeros(2 * Ip, D, device = 'cuda'); bp = torch.zeros(2 * Ip, device = 'cuda')
    col = torch.arange(Ip, device = 'cuda')
    valid = col < inner
    tile, j = col // 64, col % 64
    Wp[(tile * 128 + j)[valid]] = W1[col[valid]]; Wp[(tile * 128 + 64 + j)[valid]] = W1[inner + col[valid]]
    bp[(tile * 128 + j)[valid]] = b1[col[valid]]; bp[(tile * 128 + 64 + j)[valid]] = b1[inner + col[valid]]
    Wp = Wp.to(BF16)
    vg = torch.zeros(M, 2 * Ip, device = 'cuda', dtype = BF16); h = torch.zeros(M, Ip, device = 'cuda', dtype = BF16)
    ops.gemm_geglu(u, D, Wp, D, bp, M, 2 * Ip, D, vg, h)
    pre = u.float() @ W1.to(BF16).float().t() + b1
    val, gate = pre[:, :inner], pre[:, inner:]
    want_h = torch.nn.functional.gelu(gate) * val
    torch.cuda.synchronize()
    assert torch.allclose(h[:, :inner].float(), want_h, atol = 6e-2, rtol = 3e-2) and (h[:, inner:] == 0).all()
    got_val = vg.float().reshape(M, Ip // 64, 2, 64)[:, :, 0].reshape(M, Ip)[:, :inner]
    got_gate = vg.float().reshape(M, Ip // 64, 2, 64)[:, :, 1].reshape(M, Ip)[:, :inner]
    assert torch.allclose(got_val, val, atol = 6e-2, rtol = 2e-2) and torch.allclose(got_gate, gate, atol = 6e-2, rtol = 2e-2)
    # backward from the SAVED bf16 pre-activations
    dh = torch.zeros(M, Ip, device = 'cuda', dtype = BF16)
    dh[:, :inner] = torch.randn(M, inner, device = 'cuda', generator = g).to(BF16)
    dvg = torch.zeros_like(vg)
    rpb = ops.lib.tfx_geglu_bwd_rows_per_block()
    nblk = (M + rpb - 1) // rpb
    part = torch.zeros(nblk, 2 * Ip, device = 'cuda')
    ops.geglu_bwd(dh, vg, dvg, M, Ip, None, None, part)
    vgf = vg.float().reshape(M, Ip // 64, 2, 64)
    v_s, g_s = vgf[:, :, 0].reshape(M, Ip).clone().requires_grad_(True), vgf[:, :, 1].reshape(M, Ip).clone().requires_grad_(True)
    (torch.nn.functional.gelu(g_s) * v_s).backward(dh.float())
    torch.cuda.synchronize()
    dv_got = dvg.float().reshape(M, Ip // 64, 2, 64)[:, :, 0].reshape(M, Ip)
    dg_got = dvg.float().reshape(M, Ip // 64, 2, 64)[:, :, 1].reshape(M, Ip)
    assert torch.allclose(dv_got, v_s.grad, atol = 3e-2, rtol = 2e-2) and torch.allclose(dg_got, g_s.grad, atol = 3e-2, rtol = 2e-2)
    colsum = part.sum(0).reshape(Ip // 64, 2, 64)
    assert torch.allclose(colsum[:, 0].reshape(Ip), dvg.float().reshape(M, Ip // 64, 2, 64)[:, :, 0].reshape(M, Ip).sum(0), atol = 0.5, rtol = 2e-2)
    assert torch.allclose(colsum[:, 1].reshape(Ip), dg_got.sum(0), atol = 0.5, rtol = 2e-2)


# ================================================================================================ backward row kernels vs autograd
def test_adaln_and_resid_backward_vs_autograd(ops):
    M, D, nc = 1500, 512, 6
    g = torch.Generator(device = 'cuda').manual_seed(7)
    x = (torch.randn(M, D, device = 'cuda', generator = g) * 2 + 0.5).requires_grad_(True)
    cond_row = torch.sort(torch.randint(-1, nc, (M,), device = 'cuda', generator = g, dtype = torch.int32)).values          # runs of equal rows, as in a packed batch
    cond_row = cond_row[torch.randperm(30, device = 'cuda', generator = g).repeat_interleave(50)[:M].argsort(stable = True)]   # ... in shuffled chunks
    W3 = 7 * D
    film = (torch.randn(nc, W3, device = 'cuda', generator = g) * 0.3).requires_grad_(True)                                   # strided table, this wrapper at column D
    gam = (torch.randn(D, device = 'cuda', generator = g) * 0.3).requires_grad_(True)
    u = torch.zeros(M, D, device = 'cuda', dtype = BF16); stats = torch.zeros(M, 2, device = 'cuda')
    ops.adaln_fwd(x.detach(), cond_row, film.detach()[:, D:], W3, gam.detach(), u, stats, M, D)
    isM = (cond_row >= 0)[:, None]
    cr = cond_row.long().clamp(min = 0)
    xh = torch.nn.functional.layer_norm(x, (D,))
    want_u = torch.where(isM, xh * (film[cr, D:2 * D] + 1) + film[cr, 2 * D:3 * D], xh * (gam + 1))
    du = torch.randn(M, D, device = 'cuda', generator = g)
    want_u.backward(du)
    dx = torch.full((M, D), 0.25, device = 'cuda')                      # accumulated into
    dfilm = torch.zeros(nc, W3, device = 'cuda'); dgam = torch.zeros(D, device = 'cuda')
    ops.adaln_bwd(du, x.detach(), stats, cond_row, film.detach()[:, D:], W3, gam.detach(), dx, dfilm[:, D:], W3, dgam, M, D)
    torch.cuda.synchronize()
    assert torch.allclose(dx - 0.25, x.grad, atol = 2e-3, rtol = 2e-3)
    assert torch.allclose(dfilm, film.grad, atol = 2e-2, rtol = 2e-3) and torch.allclose(dgam, gam.grad, atol = 2e-2, rtol = 2e-3)
    # ---- output gate backward: x_out = x_res + y * s,  s = isM ? zgate[cr] : layerscale + 1
    y = torch.randn(M, D, device = 'cuda', generator = g).to(BF16)
    zg = torch.rand(nc, 3 * D, device = 'cuda', generator = g).requires_grad_(True)
    ls = (torch.randn(D, device = 'cuda', generator = g) * 0.3).requires_grad_(True)
    yf = y.float().requires_grad_(True)
    s = torch.where(isM, zg[cr, D:2 * D], ls + 1.)
    dxo = torch.randn(M, D, device = 'cuda', generator = g)
    (yf * s).backward(dxo)
    dy = torch.zeros(M, D, device = 'cuda', dtype = BF16)
    dzg = torch.zeros(nc, 3 * D, device = 'cuda'); dls = torch.zeros(D, device = 'cuda'); dbias = torch.zeros(D, device = 'cuda')
    ops.resid_bwd(dxo, y, cond_row, zg.detach()[:, D:], 3 * D, ls.detach(), dy, dzg[:, D:], 3 * D, dls, dbias, M, D)
    torch.cuda.synchronize()
    assert torch.allclose(dy.float(), yf.grad, atol = 3e-2, rtol = 2e-2)
    assert torch.allclose(dzg, zg.grad, atol = 3e-2, rtol = 3e-3) and torch.allclose(dls, ls.grad, atol = 3e-2, rtol = 3e-3)
    assert torch.allclose(dbias, dy.float().sum(0), atol = 0.2, rtol = 1e-2)


def test_attn_residual_rmsnorm_embed_backward_vs_autograd(ops):
    import ctypes
    M, D, L1 = 900, 512, 5
    g = torch.Generator(device = 'cuda').manual_seed(8)
    hid = [torch.randn(M, D, device = 'cuda', generator = g).requires_grad_(True) for _ in range(L1)]
    gam = (torch.randn(D, device = 'cuda', generator = g) * 0.3).requires_grad_(True)
    pq = (torch.randn(D, device = 'cuda', generator = g) * 0.5).requires_grad_(True)
    vals = torch.stack(hid)
    keys = torch.nn.functional.normalize(vals, dim = -1) * D ** 0.5 * (gam + 1)
    sim = torch.einsum('lnd,d->nl', keys, pq) * D ** -0.5
    want = torch.einsum('nl,lnd->nd', sim.softmax(-1), vals)
    dxo = torch.randn(M, D, device = 'cuda', generator = g)
    want.backward(dxo)
    parr = lambda ts: ctypes.cast((ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts]), ctypes.c_void_p)
    hd = [h.detach() for h in hid]
    xo = torch.zeros(M, D, device = 'cuda'); lse = torch.zeros(M, device = 'cuda')
    keep1 = (ctypes.c_void_p * L1)(*[t.data_ptr() for t in hd])
    ops.attn_residual_fwd(ctypes.cast(keep1, ctypes.c_void_p), L1, gam.detach(), pq.detach(), xo, None, lse, M, D)
    dh = [torch.full((M, D), 0.5, device = 'cuda') for _ in range(L1)]
    keep2 = (ctypes.c_void_p * L1)(*[t.data_ptr() for t in dh])
    dgam = torch.zeros(D, device = 'cuda'); dpq = torch.zeros(D, device = 'cuda')
    ws = torch.zeros(int(ops.lib.tfx_attn_residual_bwd_workspace_floats(M, D)), device = 'cuda')
    ops.attn_residual_bwd(ctypes.cast(keep1, ctypes.c_void_p), ctypes.cast(keep2, ctypes.c_void_p), L1, gam.detach(), pq.detach(), dxo, xo, lse, dgam, dpq, ws, M, D, 0)
    torch.cuda.synchronize()
    assert torch.allclose(xo, want.detach(), atol = 1e-4, rtol = 1e-4)
    for l in range(L1):
        assert torch.allclose(dh[l] - 0.5, hid[l].grad, atol = 2e-4, rtol = 2e-3), l
    assert torch.allclose(dgam, gam.grad, atol = 2e-3, rtol = 5e-3) and torch.allclose(dpq, pq.grad, atol = 2e-3, rtol = 5e-3)
    # init = 1 overwrites the hidden gradients
    ops.attn_residual_bwd(ctypes.cast(keep1, ctypes.c_void_p), ctypes.cast(keep2, ctypes.c_void_p), L1, gam.detach(), pq.detach(), dxo, xo, lse, dgam, dpq, ws, M, D, 1)
    torch.cuda.synchronize()
    assert torch.allclose(dh[2], hid[2].grad, atol = 2e-4, rtol = 2e-3)
    # ---- final RMSNorm
    x = (torch.randn(M, D, device = 'cuda', generator = g) * 1.5).requires_grad_(True)
    gn = (torch.randn(D, device = 'cuda', generator = g) * 0.3).requires_grad_(True)
    out = torch.nn.functional.normalize(x, dim = -1) * D ** 0.5 * (gn + 1)
    dout = torch.randn(M, D, device = 'cuda', generator = g)
    out.backward(dout)
    of = torch.zeros(M, D, device = 'cuda'); ob = torch.zeros(M, D, device = 'cuda', dtype = BF16)
    ops.rmsnorm_fwd(x.detach(), gn.detach(), of, ob, None, None, M, D)
    dx = torch.zeros(M, D, device = 'cuda'); dgn = torch.zeros(D, device = 'cuda')
    ops.rmsnorm_bwd(dout, x.detach(), gn.detach(), dx, dgn, M, D)
    torch.cuda.synchronize()
    assert torch.allclose(of, out.detach(), atol = 1e-4, rtol = 1e-4)
    assert torch.allclose(dx, x.grad, atol = 2e-4, rtol = 2e-3) and torch.allclose(dgn, gn.grad, atol = 5e-3, rtol = 5e-3)
    # ---- token assemble backward: text rows scatter-add into the embedding gradient, modality rows go to the compact matrix
    V, S = 70, 300
    text_id = torch.randint(0, V, (M,), device = 'cuda', generator = g, dtype = torch.int32)
    slot = torch.full((M,), -1, device = 'cuda', dtype = torch.int32)
    rows = torch.randperm(M, device = 'cuda', generator = g)[:S]
    slot[rows] = torch.arange(S, device = 'cuda', dtype = torch.int32)
    dx0 = torch.randn(M, D, device = 'cuda', generator = g)
    demb = torch.zeros(V, D, device = 'cuda'); dmod = torch.zeros(S, D, device = 'cuda', dtype = BF16)
    ops.embed_bwd(dx0, text_id, slot, demb, dmod, M, D)
    want_emb = torch.zeros(V, D, device = 'cuda')
    is_text = slot < 0
    want_emb.index_add_(0, text_id[is_text].long(), dx0[is_text])
    torch.cuda.synchronize()
    assert torch.allclose(demb, want_emb, atol = 1e-3, rtol = 1e-4)
    assert torch.allclose(dmod.float(), dx0[rows], atol = 3e-2, rtol = 1e-2)


def test_qk_bwd_pack_vs_autograd(ops):
    """backward of the qk-RMSNorm + RoPE epilogue (T.py:950-965) and of the value gate logits (T.py:1026-1027)"""
    M, H = 640, 4
    HI, NQ = H * 64, 3 * H * 64 + 128
    g = torch.Generator(device = 'cuda').manual_seed(9)
    freqs, t, tt = _rope_tables(ops, 1024)
    pos = torch.randint(0, 900, (M,), device = 'cuda', generator = g, dtype = torch.int32)
    xq = torch.randn(M, H, 64, device = 'cuda', generator = g).requires_grad_(True)
    xk = torch.randn(M, H, 64, device = 'cuda', generator = g).requires_grad_(True)
    gq = (torch.randn(64, device = 'cuda', generator = g) * 0.3).requires_grad_(True)
    gk = (torch.randn(64, device = 'cuda', generator = g) * 0.3).requires_grad_(True)
    rms = lambda x, gm: torch.nn.functional.normalize(x, dim = -1) * 8. * (gm + 1.)
    q = _rope_ref(rms(xq, gq), pos, freqs); k = _rope_ref(rms(xk, gk), pos, freqs)
    dq = torch.randn(M, HI, device = 'cuda', generator = g); dk = torch.randn(M, HI, device = 'cuda', generator = g)
    (q.reshape(M, HI) * dq).sum().backward(retain_graph = True)
    (k.reshape(M, HI) * dk).sum().backward()
    inv = torch.cat((1. / xq.detach().norm(dim = -1), 1. / xk.detach().norm(dim = -1)), dim = 1).contiguous()          # [M, 2H]
    gates = torch.randn(M, H, device = 'cuda', generator = g)
    dsum = torch.randn(M, H, device = 'cuda', generator = g)
    out = torch.zeros(M, NQ, device = 'cuda', dtype = BF16)
    dgq = torch.zeros(64, device = 'cuda'); dgk = torch.zeros(64, device = 'cuda')
    ops.qk_bwd_pack(dq, dk, q.detach().reshape(M, HI).to(BF16), k.detach().reshape(M, HI).to(BF16), inv, gq.detach(), gk.detach(), pos, t, gates, dsum, out, NQ, dgq, dgk, M, H)
    torch.cuda.synchronize()
    # the kernel reconstructs xhat from the bf16 q / k it is given: tolerances are those of bf16 inputs
    assert torch.allclose(out[:, :HI].float(), xq.grad.reshape(M, HI), atol = 8e-2, rtol = 5e-2)
    assert torch.allclose(out[:, HI:2 * HI].float(), xk.grad.reshape(M, HI), atol = 8e-2, rtol = 5e-2)
    assert torch.allclose(out[:, 3 * HI:3 * HI + H].float(), (1 - torch.sigmoid(gates)) * dsum, atol = 2e-2, rtol = 2e-2)
    assert torch.allclose(dgq, gq.grad, atol = 0.5, rtol = 3e-2) and torch.allclose(dgk, gk.grad, atol = 0.5, rtol = 3e-2)


# ================================================================================================ decode-path kernels
def test_attn_decode_vs_dense(ops):
    """single-query decode attention over cache slabs (csrc/decode.cu) against the dense formula, incl. slabs of very different fill"""
    H, cap_rows, S = 4, 700, 5
    scale, softcap = 0.125, 50.
    g = torch.Generator(device = 'cuda').manual_seed(10)
    lens = [1, 33, 128, 300, 699]
    kc = (torch.randn(S * cap_rows, H * 64, device = 'cuda', generator = g) * 1.5).to(BF16); vc = (torch.randn(S * cap_rows, H * 64, device = 'cuda', generator = g) * 2).to(BF16)
    q = (torch.randn(S, H * 64, device = 'cuda', generator = g) * 1.5).to(BF16)
    gates = torch.randn(S, H, device = 'cuda', generator = g)
    i32 = lambda v: torch.tensor(v, device = 'cuda', dtype = torch.int32)
    kv0 = i32([s * cap_rows for s in range(S)]); kvend = i32([s * cap_rows + lens[s] for s in range(S)])
    lim = kvend - 1
    o = torch.zeros(S, H * 64, device = 'cuda', dtype = BF16)
    ops.attn_decode(q, kc, vc, H * 64, H * 64, H * 64, gates, H, lim, i32(list(range(S))), kv0, kvend, S, o, H * 64, scale, softcap)
    torch.cuda.synchronize()
    for s in range(S):
        kk = kc[s * cap_rows: s * cap_rows + lens[s]].float().reshape(lens[s], H, 64)
        vv = vc[s * cap_rows: s * cap_rows + lens[s]].float().reshape(lens[s], H, 64)
        sim = torch.einsum('hd,jhd->hj', q[s].float().reshape(H, 64) * scale, kk)
        sim = torch.tanh(sim / softcap) * softcap
        want = torch.einsum('hj,jhd->hd', sim.softmax(-1), vv) * torch.sigmoid(gates[s])[:, None]
        assert torch.allclose(o[s].float().reshape(H, 64), want, atol = 3e-2, rtol = 3e-2), s


def test_sample_tokens_and_decode_prep(ops):
    """tfx_sample_tokens: greedy = torch.argmax; min-p + Gumbel draws follow the filtered softmax; the state machine of T.py:2330-2349.
    tfx_decode_prep: state -> metadata of the next text step."""
    S, V, ld = 64, 390, 392
    g = torch.Generator(device = 'cuda').manual_seed(11)
    logits = torch.randn(S, ld, device = 'cuda', generator = g) * 2
    st = torch.zeros(6, S, device = 'cuda', dtype = torch.int32)
    st[0] = 10; st[1] = 7
    st[3, 5] = 2; st[3, 6] = 1                                          # one finished sample, one waiting for its modality: untouched
    hist = torch.zeros(S, 8, device = 'cuda', dtype = torch.int32); cnt = torch.zeros(2, device = 'cuda', dtype = torch.int32)
    som = torch.tensor([259], device = 'cuda', dtype = torch.int32)
    logits[3, 257] = 50.; logits[4, 259] = 50.                          # sample 3 draws [eos], sample 4 draws [som]
    before = st.clone()
    ops.sample_tokens(logits, ld, None, V, 0, st, S, hist, 8, 257, som, 1, 1000, 0.0, 0.1, 1234, cnt, 1)
    torch.cuda.synchronize()
    want = logits[:, :V].argmax(-1).int()
    act = torch.ones(S, dtype = torch.bool, device = 'cuda'); act[5] = act[6] = False
    assert torch.equal(st[2][act], want[act]) and torch.equal(hist[:, 0][act], want[act]) and torch.equal(st[:, ~act], before[:, ~act])
    assert (st[0][act] == 11).all() and (st[1][act] == 8).all() and (st[4][act] == 1).all() and (st[5][act] == 1).all()
    assert st[3, 3].item() == 2 and st[3, 4].item() == 1 and cnt[0].item() == int(act.sum()) - 2
    # length limit: num_tokens > max_length ends the sample
    st2 = torch.zeros(6, S, device = 'cuda', dtype = torch.int32); st2[4] = 5
    ops.sample_tokens(logits, ld, None, V, 0, st2, S, hist, 8, -1, som, 0, 5, 0.0, 0.1, 1, cnt, 0)
    torch.cuda.synchronize()
    assert (st2[3] == 2).all() and (st2[0] == 0).all() and (st2[1] == 0).all()          # advance = 0: first token after the prefill
    # Gumbel-max draws: empirical frequencies over many (sample, step) pairs follow softmax(min-p filtered logits / T), restricted to ids < vlimit
    Vs, T, minp = 12, 0.7, 0.2
    base = torch.tensor([2.0, 1.5, 1.0, 0.0, -1.0, -3.0, 0.5, 1.8, -0.5, 0.2, 3.0, 2.5], device = 'cuda')
    lg = torch.zeros(4096, 16, device = 'cuda'); lg[:, :Vs] = base
    counts = torch.zeros(Vs, device = 'cuda')
    for step in range(4):
        st3 = torch.zeros(6, 4096, device = 'cuda', dtype = torch.int32); h3 = torch.zeros(4096, 2, device = 'cuda', dtype = torch.int32)
        c3 = torch.tensor([0, step], device = 'cuda', dtype = torch.int32)
        ops.sample_tokens(lg, 16, None, Vs, 10, st3, 4096, h3, 2, -1, som, 0, 10 ** 6, T, minp, 777, c3, 1)
        counts += torch.bincount(st3[2].long(), minlength = Vs).float()[:Vs]
    torch.cuda.synchronize()
    x = base / T
    p = x.softmax(-1)
    keep = p >= minp * p.max()                                          # min-p over ALL logits (T.py:574-578) ...
    keep[10:] = False                                                   # ... then the text-only restriction (T.py:2697): ids 10, 11 carry the largest logits
    want_p = torch.where(keep, p, torch.zeros_like(p)); want_p = want_p / want_p.sum()
    freq = counts / counts.sum()
    assert (counts[~keep] == 0).all()
    assert (freq - want_p).abs().max().item() < 0.02, (freq, want_p)
    # decode_prep
    st4 = torch.zeros(6, 4, device = 'cuda', dtype = torch.int32)
    st4[0] = torch.tensor([3, 0, 99, 50]); st4[1] = torch.tensor([2, 0, 40, 7]); st4[2] = torch.tensor([11, 12, 13, 14])
    meta = torch.zeros(8, 4, device = 'cuda', dtype = torch.int32); c4 = torch.zeros(2, device = 'cuda', dtype = torch.int32)
    ops.decode_prep(st4, 4, 100, 2, meta[0], meta[1], meta[2], meta[3], meta[4], meta[5], meta[6], meta[7], c4)
    torch.cuda.synchronize()
    base_rows = torch.tensor([200, 300, 400, 500], device = 'cuda', dtype = torch.int32)
    assert torch.equal(meta[0], st4[2]) and torch.equal(meta[1], st4[1]) and torch.equal(meta[2], base_rows + st4[0]) and torch.equal(meta[3], meta[2])
    assert meta[4].tolist() == [0, 1, 2, 3] and meta[5].tolist() == [1, 2, 3, 4] and torch.equal(meta[6], base_rows) and torch.equal(meta[7], base_rows + st4[0] + 1)
    assert c4.tolist() == [0, 1]


def test_ode_kernels_follow_the_midpoint_rule(ops):
    """tfx_ode_pre / tfx_ode_post over a table from decode.midpoint_table integrate dy/dt = a(t) y exactly like the host-side midpoint loop"""
    from transfusion_pytorch_b200.decode import midpoint_table
    steps, n = 6, 1000
    tab = midpoint_table(steps, 'cuda')
    y = torch.randn(n, device = 'cuda'); y0 = y.clone()
    fprev = torch.zeros(n, device = 'cuda'); x = torch.zeros(2 * n, device = 'cuda'); ct = torch.zeros(3, device = 'cuda')
    idx = torch.zeros(1, device = 'cuda', dtype = torch.int32)
    f = lambda t, v: (0.5 - t) * v + 0.1
    for e in range(2 * (steps - 1)):
        ops.ode_pre(y, fprev, x, n, 2, tab, idx, ct, 3)
        t = ct[0].item()
        assert torch.equal(x[:n], x[n:]) and (ct == ct[0]).all()
        pc, pu = f(t, x[:n]) * 1.5, f(t, x[:n]) * 0.5                   # cfg 2: u + 2 (c - u) = 2.5 f ... use cfg = 0.5: u + 0.5 (c - u) = f
        ops.ode_post(y, fprev, pc, pu, 0.5, n, tab, idx)
        ops.counter_inc(idx)
    grid = torch.linspace(0, 1, steps)
    w = y0.clone()
    for t0, t1 in zip(grid[:-1], grid[1:]):
        dt = (t1 - t0).item()
        k1 = f(t0.item(), w)
        w = w + dt * f(t0.item() + 0.5 * dt, w + 0.5 * dt * k1)
    torch.cuda.synchronize()
    assert torch.allclose(y, w, atol = 1e-5, rtol = 1e-5)

def test_attn_residual_deferred_backward_chain_vs_autograd(ops):
    """tfx_attn_residual_bwd2: a stack of 4 AttentionResiduals over 5 hiddens (layer i mixes h_0..h_{i+1}), loss = sum_i <x_i, R_i>.  The deferred kernels
    assemble the COMPLETE gradient of each hidden once (own layer + stored scalars of the later layers) and must equal autograd's sum over layers."""
    import ctypes
    M, D, depth = 700, 512, 4
    g = torch.Generator(device = 'cuda').manual_seed(11)
    hid = [torch.randn(M, D, device = 'cuda', generator = g).to(BF16).float().requires_grad_(True) for _ in range(depth + 1)]
    gams = [(torch.randn(D, device = 'cuda', generator = g) * 0.3).requires_grad_(True) for _ in range(depth)]
    pqs = [(torch.randn(D, device = 'cuda', generator = g) * 0.5).requires_grad_(True) for _ in range(depth)]
    R = [torch.randn(M, D, device = 'cuda', generator = g) for _ in range(depth)]
    loss = 0.
    for i in range(depth):
        vals = torch.stack(hid[:i + 2])
        keys = torch.nn.functional.normalize(vals, dim = -1) * D ** 0.5 * (gams[i] + 1)
        sim = torch.einsum('lnd,d->nl', keys, pqs[i]) * D ** -0.5
        loss = loss + (torch.einsum('nl,lnd->nd', sim.softmax(-1), vals) * R[i]).sum()
    loss.backward()
    hb = [h.detach().to(BF16) for h in hid]
    keep = []
    def parr(ts):
        a = (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts]); keep.append(a)
        return ctypes.cast(a, ctypes.c_void_p)
    xo = [torch.zeros(M, D, device = 'cuda') for _ in range(depth)]; lse = [torch.zeros(M, device = 'cuda') for _ in range(depth)]
    for i in range(depth):
        ops.attn_residual_fwd_h16(parr(hb[:i + 2]), i + 2, gams[i].detach(), pqs[i].detach(), xo[i], None, lse[i], M, D)
    stride = (depth + 2) * 3
    sc = torch.zeros(depth, M, depth + 2, 3, device = 'cuda')
    G = [torch.full((M, D), 7., device = 'cuda') for _ in range(depth + 1)]
    dgam = [torch.zeros(D, device = 'cuda') for _ in range(depth)]; dpq = [torch.zeros(D, device = 'cuda') for _ in range(depth)]
    ws = torch.zeros(int(ops.lib.tfx_attn_residual_bwd_workspace_floats(M, D)), device = 'cuda')
    gd, pd = [t.detach() for t in gams], [t.detach() for t in pqs]
    for i in reversed(range(depth)):
        later = list(range(i + 1, depth))
        ops.attn_residual_bwd2(parr(hb[:i + 2]), i + 2, 1, parr([gd[j] for j in [i] + later]), parr([pd[j] for j in [i] + later]), parr([R[j] for j in later] or [R[i]]),
                               parr([sc[j][0, i + 1] for j in later] or [R[i]]), len(later), R[i], xo[i], lse[i], G[i + 1], sc[i], stride, dgam[i], dpq[i], ws, M, D)
    allj = list(range(depth))
    ops.attn_residual_bwd2(parr(hb[:1]), 1, 0, parr([gd[0]] + gd), parr([pd[0]] + pd), parr(R), parr([sc[j][0, 0] for j in allj]), depth, None, None, None, G[0], None, stride, None, None, None, M, D)
    torch.cuda.synchronize()
    for k in range(depth + 1):
        err = (G[k] - hid[k].grad).abs().max().item() / hid[k].grad.abs().max().item()
        assert err < 2e-3, (k, err)
    for i in range(depth):
        assert torch.allclose(dgam[i], gams[i].grad, atol = 5e-3, rtol = 1e-2) and torch.allclose(dpq[i], pqs[i].grad, atol = 5e-3, rtol = 1e-2), i
