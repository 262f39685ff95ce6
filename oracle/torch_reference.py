"""TEST INFRASTRUCTURE ONLY - plain-PyTorch fp32 restatement of the reference's hot-path ALGORITHM
(lucidrains/transfusion-pytorch v0.19.4), written against the ragged descriptor of
transfusion_pytorch_b200.modality_processing so it can stand in for the CUDA engine in CPU tests of the
host logic (`OracleEngine`) and be timed as the CPU baseline (`bench.py --impl reference`, kind "port").

Pinned: tests/test_oracle_cpu.py checks it against tests/golden/*.pt, which are outputs of the reference
itself (oracle/make_golden.py).  It mirrors the reference's COST structure on purpose - conditioning
evaluated per token (T.py:1132, 749, 767), dense N x N scores with an explicit boolean mask
(T.py:452-470, 998-1013), padded batches - so that timing it is a fair stand-in for the reference's CPU path.

Parity unpinned UPSTREAM for two conventions only: the RoPE pairing / base (third-party `rotary_embedding_torch>=0.8.4`) and the
fixed-grid midpoint solver (`torchdiffeq`) are not in /root/reference; their published algorithms are restated in oracle/shims and the
reference's own self-consistency tests (tests/test_transfusion.py:559-662, 758-808 of the reference) pass through them.

Citations are to /root/reference/transfusion_pytorch/transfusion.py ("T.py").
Never imported by the product package.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _rms(x, gamma):                                   # T.py:779-786
    return F.normalize(x, dim = -1) * (x.shape[-1] ** 0.5) * (gamma + 1.)


def _rope(x, pos, freqs):                             # rotary_embedding_torch, interleaved pairs (T.py:965)
    ang = pos[..., None].float() * freqs              # [..., n, 32]
    ang = ang.repeat_interleave(2, dim = -1)
    x2 = x.reshape(*x.shape[:-1], -1, 2)
    rot = torch.stack((-x2[..., 1], x2[..., 0]), dim = -1).flatten(-2)
    return x * ang.cos() + rot * ang.sin()


class TorchReference:
    """Forward (with autograd) of the block stack + heads on padded [b, n] batches rebuilt from a RaggedBatch."""

    def __init__(self, model):
        self.m = model
        self.tr = model.transformer

    def P(self, name):
        return dict(self.m.named_parameters())[name]

    def padded(self, rb):
        B, n = rb.B, int(rb.seq_lens.max())
        idx = torch.full((B, n), -1, dtype = torch.long)
        for b in range(B):
            idx[b, :rb.seq_lens[b]] = torch.arange(rb.cu[b], rb.cu[b + 1])
        return idx

    def stack(self, rb, x0, cond_tok, is_mod, kv_limit, rope_pos, valid, attend = None):
        """x0 [b,n,d]; cond_tok [b,n] time per token or None; returns final-norm output and hiddens.
        attend(layer, q, k, v) -> o overrides the dense masked attention (kv-cache decoding: T.py:969-977)."""
        tr, D, H = self.tr, self.tr.dim, self.tr.heads
        sd = dict(self.m.named_parameters())
        B, n, _ = x0.shape
        cond = None
        if cond_tok is not None:                      # per-token time conditioning (T.py:1128-1132)
            w = self.tr.to_time_cond[0].weights
            fr = cond_tok[..., None] * w * 2 * math.pi
            feats = torch.cat((cond_tok[..., None], fr.sin(), fr.cos()), dim = -1)
            cond = F.silu(F.linear(feats, sd['transformer.to_time_cond.1.weight'], sd['transformer.to_time_cond.1.bias']))
        j = torch.arange(n)
        mask = (j[None, None, :] <= kv_limit[:, :, None]) & valid[:, None, :]          # T.py:452-470 as j <= kv_limit[i]
        freqs = self.m.rotary_emb.freqs
        x, hid, skips = x0, [x0], []
        isM = is_mod[..., None]

        def wrap_in(x, pre):
            xh = F.layer_norm(x, (D,))
            t = xh * (sd[f'{pre}.layernorm_gamma'] + 1.)
            if cond is None:
                return t
            g, b = F.linear(cond, sd[f'{pre}.to_film.weight'], sd[f'{pre}.to_film.bias']).chunk(2, dim = -1)
            return torch.where(isM, xh * (g + 1.) + b, t)                                 # T.py:747-755

        def wrap_out(y, pre):
            t = y * (sd[f'{pre}.layerscale'] + 1.)
            if cond is None:
                return t
            z = F.linear(cond, sd[f'{pre}.to_ada_ln_zero.weight'], sd[f'{pre}.to_ada_ln_zero.bias']).sigmoid()
            return torch.where(isM, y * z, t)                                             # T.py:765-769

        for i in range(tr.depth):
            pre = f'transformer.layers.{i}'
            layer = i + 1
            if layer <= tr.depth // 2:
                skips.append(x)
            elif f'{pre}.0.weight' in sd:
                x = F.linear(torch.cat((x, skips.pop()), dim = -1), sd[f'{pre}.0.weight']) + x   # T.py:1214-1219
            u = wrap_in(x, f'{pre}.1')
            qk = F.linear(u, sd[f'{pre}.1.fn.to_qk.0.weight']).reshape(B, n, 2, H, 64)
            q, k = qk[:, :, 0].transpose(1, 2), qk[:, :, 1].transpose(1, 2)
            v = F.linear(u, sd[f'{pre}.1.fn.to_v.0.weight']).reshape(B, n, H, 64).transpose(1, 2)
            q, k = _rms(q, sd[f'{pre}.1.fn.q_norm.gamma']), _rms(k, sd[f'{pre}.1.fn.k_norm.gamma'])
            q, k = _rope(q, rope_pos[:, None], freqs), _rope(k, rope_pos[:, None], freqs)
            cap = tr.softcap_value
            if attend is not None:
                o = attend(i, q, k, v)
            else:
                sim = torch.einsum('bhid,bhjd->bhij', q * 64 ** -0.5, k)
                sim = (sim / cap).tanh() * cap                                              # T.py:1001
                sim = sim.masked_fill(~mask[:, None], -torch.finfo(sim.dtype).max)
                o = torch.einsum('bhij,bhjd->bhid', sim.softmax(dim = -1), v)
            o = o * F.linear(u, sd[f'{pre}.1.fn.to_gates.0.weight']).transpose(1, 2)[..., None].sigmoid()   # T.py:1026-1027
            a = F.linear(o.transpose(1, 2).reshape(B, n, H * 64), sd[f'{pre}.1.fn.to_out.1.weight'])
            x = x + wrap_out(a, f'{pre}.1')
            u = wrap_in(x, f'{pre}.2')
            hcat = F.linear(u, sd[f'{pre}.2.fn.net.0.weight'], sd[f'{pre}.2.fn.net.0.bias'])
            val, gate = hcat.chunk(2, dim = -1)
            f = F.linear(F.gelu(gate) * val, sd[f'{pre}.2.fn.net.3.weight'], sd[f'{pre}.2.fn.net.3.bias'])   # T.py:833-834
            x = x + wrap_out(f, f'{pre}.2')
            hid.append(x)
            vals = torch.stack(hid)                                                         # AttentionResidual T.py:803-829
            keys = _rms(vals, sd[f'{pre}.3.norm_keys.gamma'])
            sim_l = torch.einsum('lbnd,d->bnl', keys, sd[f'{pre}.3.pseudo_queries']) * D ** -0.5
            x = torch.einsum('bnl,lbnd->bnd', sim_l.softmax(dim = -1), vals)
        out = _rms(x, sd['transformer.norm.gamma'])
        return out, hid

    def cached_attention(self, rb, cache):
        """attention of NEW tokens against a slab kv cache: append the new keys / values at rb.kv_row (T.py:969-972), attend rows
        [slab start, kv_limit[i]] of the sample's slab (the explicit masks of T.py:2300-2304, 2415-2431 in closed form)."""
        cap_rows, softcap = cache.cap, self.tr.softcap_value
        def attend(layer, q, k, v):
            o = torch.zeros_like(q)
            for b in range(rb.B):
                s0, n = int(rb.cu[b]), int(rb.seq_lens[b])
                if n == 0:
                    continue
                rows = torch.as_tensor(rb.kv_row[s0:s0 + n]).long()
                cache.k[layer][rows] = k[b, :, :n].transpose(0, 1)
                cache.v[layer][rows] = v[b, :, :n].transpose(0, 1)
                start = (int(rows[0]) // cap_rows) * cap_rows
                end = int(rb.kv_limit[s0:s0 + n].max()) + 1
                kk, vv = cache.k[layer][start:end].transpose(0, 1), cache.v[layer][start:end].transpose(0, 1)      # [H, L, 64]
                sim = torch.einsum('hid,hjd->hij', q[b, :, :n] * 64 ** -0.5, kk)
                sim = (sim / softcap).tanh() * softcap
                j = torch.arange(start, end)
                vis = j[None, :] <= torch.as_tensor(rb.kv_limit[s0:s0 + n]).long()[:, None]
                sim = sim.masked_fill(~vis[None], -torch.finfo(sim.dtype).max)
                o[b, :, :n] = torch.einsum('hij,hjd->hid', sim.softmax(dim = -1), vv)
            return o
        return attend

    def run(self, rb, latents, eps, *, text_loss_weight = 1., flow_loss_weight = 1., vlimit = 0, modality_only = False, want_loss = True, cache = None):
        m, D = self.m, self.tr.dim
        sd = dict(m.named_parameters())
        idx = self.padded(rb)
        valid = idx >= 0
        gi = idx.clamp(min = 0)
        tens = lambda a, dt = torch.long: torch.as_tensor(a).to(dt)
        text_id, label = tens(rb.text_id)[gi], tens(rb.label)[gi].masked_fill(~valid, -1)
        B, n = idx.shape
        seq_start = torch.as_tensor(rb.cu[:-1])[:, None]
        kv_limit = (tens(rb.kv_limit)[gi] - seq_start).masked_fill(~valid, 0)
        kv_limit = torch.where(valid, kv_limit, torch.arange(n)[None].expand(B, n))
        rope_pos = tens(rb.rope_pos)[gi]
        cond_row, slot = tens(rb.cond_row)[gi].masked_fill(~valid, -1), tens(rb.slot)[gi].masked_fill(~valid, -1)
        is_mod = cond_row >= 0
        x0 = m.text_embed.weight[text_id]
        flows, noised_all = [None] * rb.n_types, None
        if rb.S > 0:
            modtok = torch.zeros(rb.S, D)
            rt = torch.as_tensor(rb.row_time)
            for t, (s0, s1) in enumerate(rb.type_rows):
                if s1 == s0:
                    continue
                x = latents[t].float().cpu()
                if eps is not None and eps[t] is not None:
                    tt = rt[s0:s1, None]
                    e = eps[t].float().cpu()
                    flows[t] = x - e                                                        # MP.py:645-656
                    x = x * tt + e * (1. - tt)
                proj = m.latent_to_model_projs[t]
                modtok[s0:s1] = proj(x) if not isinstance(proj, torch.nn.Identity) else x
            x0 = torch.where(is_mod[..., None], modtok[slot.clamp(min = 0)], x0)            # T.py:3184
        cond_tok = None
        if rb.n_cond > 0:
            ct = torch.as_tensor(rb.cond_times)
            cond_tok = torch.where(is_mod, ct[cond_row.clamp(min = 0)], torch.zeros(()))    # T.py:3230-3232
        out, hid = self.stack(rb, x0, cond_tok, is_mod, kv_limit, rope_pos, valid, attend = self.cached_attention(rb, cache) if cache is not None else None)
        res = dict(embed = out, hiddens = hid, valid = valid)
        logits = F.linear(out, sd['to_text_logits.weight'])
        res['logits'] = logits
        preds = [None] * rb.n_types
        for t, (s0, s1) in enumerate(rb.type_rows):
            if s1 > s0:
                rows = torch.as_tensor(rb.row_token[s0:s1]).long()
                b_of = torch.searchsorted(torch.as_tensor(rb.cu[1:]), rows, right = True)
                pos = rows - torch.as_tensor(rb.cu)[b_of]
                preds[t] = F.linear(out[b_of, pos], sd[f'model_to_latent_projs.{t}.weight'])  # T.py:3301-3302
        res['preds'] = preds
        if not want_loss:
            return res
        T = float(rb.total_tokens)
        if vlimit:
            lg = logits.masked_fill(~(torch.arange(logits.shape[-1]) < vlimit), -torch.finfo(logits.dtype).max)   # T.py:2653
            res['total'] = res['text'] = F.cross_entropy(lg.transpose(1, 2), label, ignore_index = -1)
            res['flows'] = torch.zeros(0)
            return res
        text = F.cross_entropy(logits.transpose(1, 2), label, ignore_index = -1) if (label >= 0).any() else logits.sum() * 0.
        fl = [F.mse_loss(preds[t], flows[t]) if flows[t] is not None else torch.zeros(()) for t in range(rb.n_types)]
        if modality_only:
            total = sum(fl)
        else:
            total = text * ((label >= 0).sum() / T) * text_loss_weight                      # T.py:3331-3376
            for t in range(rb.n_types):
                if flows[t] is not None:
                    total = total + fl[t] * (rb.n_type_tokens[t] / T) * flow_loss_weight
        res.update(total = total, text = text, flows = torch.stack(fl) if fl else torch.zeros(0))
        return res


class OracleEngine:
    """Drop-in for transfusion_pytorch_b200.engine.Engine in CPU tests of the host logic (injected by the test:
    `model._engine = OracleEngine(model)`).  The product never constructs it."""

    def __init__(self, model):
        if any(getattr(model, 'add_pos_emb', ())):
            raise NotImplementedError('the CPU checker does not restate the axial positional embedding: that feature is pinned by tests/golden/small_posemb.pt (reference + shim)')
        self.ref = TorchReference(model)
        self.model = model
        self.state = None

    frozen = False

    def pack_weights(self, force = False):
        pass

    def forward(self, rb, latents, eps, *, train, want_logits = False, vlimit = 0, text_loss_weight = 1., flow_loss_weight = 1., modality_only = False, cache = None,
                want_preds = None):
        with torch.set_grad_enabled(train):
            res = self.ref.run(rb, latents, eps, text_loss_weight = text_loss_weight, flow_loss_weight = flow_loss_weight, vlimit = vlimit,
                               modality_only = modality_only, want_loss = train, cache = cache)
        valid = res['valid']
        pack = lambda t: t[valid]
        out = dict(embed = pack(res['embed']).detach(), logits = pack(res['logits']).detach(), preds = [p.detach() if p is not None else None for p in res['preds']])
        if train:
            self.state = res
            out.update(total = res['total'].detach(), text = res['text'].detach(), flows = res['flows'].detach())
        return out

    def backward(self, gscale = None, bucket_cb = None):
        total = self.state['total']
        total.backward(gscale.detach().cpu() if gscale is not None else None)


# ------------------------------------------------------------------------------------------------ kv-cache decode doubles (tests only)
class OracleKVCache:
    """fp32 stand-in for transfusion_pytorch_b200.engine.KVCache: same slab addressing"""
    def __init__(self, model, n_slabs, cap):
        tr = model.transformer
        self.n_slabs, self.cap, self.rows = n_slabs, cap, n_slabs * cap
        self.k = torch.zeros(tr.depth, self.rows, tr.heads, 64)
        self.v = torch.zeros(tr.depth, self.rows, tr.heads, 64)


class OracleTextDecoder:
    """Python restatement of the device-side text loop (csrc/decode.cu: tfx_decode_prep + tfx_sample_tokens around one incremental forward;
    reference step_text T.py:2279-2349).  Same interface as transfusion_pytorch_b200.decode.TextDecoder."""
    def __init__(self, engine, cache, S, *, slab0 = 0, hist_cap, eos_id, som_ids, max_length, temperature, min_p, vlimit = 0, seed = 0, use_graph = True, poll = 8):
        import numpy as np
        self.np = np
        self.eng, self.cache, self.S, self.slab0 = engine, cache, S, slab0
        self.eos_id, self.som_ids, self.max_length = eos_id, list(som_ids or []), max_length
        self.temperature, self.min_p, self.vlimit = temperature, min_p, vlimit
        self.gen = torch.Generator().manual_seed(int(seed))
        self.state = np.zeros((6, S), dtype = np.int64)
        self.hist = [[] for _ in range(S)]
        self.text_left = 0

    def set_state(self, length, tokens_seen, last_token, phase, num_tokens):
        for r, a in enumerate((length, tokens_seen, last_token, phase, num_tokens)):
            self.state[r] = self.np.asarray(a, dtype = self.np.int64)
        self.state[5] = 0
        self.hist = [[] for _ in range(self.S)]

    def update_rows(self, rows):
        for r, a in rows.items():
            self.state[r] = self.np.asarray(a, dtype = self.np.int64)
        self.state[5] = 0
        self.hist = [[] for _ in range(self.S)]

    def get_state(self):
        return self.state.copy(), [self.np.asarray(h, dtype = self.np.int64) for h in self.hist]

    def _pick(self, logits):
        if self.temperature == 0.:
            return int(logits.argmax())
        x = logits / self.temperature
        probs = x.softmax(dim = -1)
        x = torch.where(probs < self.min_p * probs.amax(), torch.tensor(float('-inf')), x)          # min_p_filter, T.py:574-578
        if self.vlimit:
            x[self.vlimit:] = float('-inf')
        u = torch.rand(x.shape, generator = self.gen).clamp(min = 1e-20)
        return int((x - torch.log(-torch.log(u))).argmax())

    def _update(self, logits_rows, advance):
        st = self.state
        left = 0
        for s in range(self.S):
            if st[3, s] != 0:
                continue
            tok = self._pick(logits_rows[s])
            self.hist[s].append(tok); st[5, s] += 1
            st[2, s] = tok
            if advance:
                st[0, s] += 1; st[1, s] += 1
            st[4, s] += 1
            if tok == self.eos_id: st[3, s] = 2
            elif st[4, s] > self.max_length: st[3, s] = 2
            elif tok in self.som_ids: st[3, s] = 1
            left += int(st[3, s] == 0)
        self.text_left = left

    def sample_first(self, logits, rows):
        V = self.eng.model.to_text_logits.weight.shape[0]
        self._update([logits[int(r), :V].float() for r in rows], advance = 0)

    def step(self):
        np, S, cap, st = self.np, self.S, self.cache.cap, self.state
        from transfusion_pytorch_b200.modality_processing import RaggedBatch
        ln = np.minimum(st[0], cap - 1)
        base = (self.slab0 + np.arange(S)) * cap
        z = np.zeros(S, dtype = np.int32)
        rb = RaggedBatch(B = S, M = S, seq_lens = np.ones(S, dtype = np.int64), cu = np.arange(S + 1, dtype = np.int64), full_lens = np.ones(S, dtype = np.int64),
                         text_id = st[2].astype(np.int32), label = z - 1, kv_limit = (base + ln).astype(np.int32), rope_pos = st[1].astype(np.int32), cond_row = z - 1, slot = z - 1,
                         n_cond = 0, cond_times = np.zeros(0, np.float32), n_types = 0, type_rows = [], row_token = np.zeros(0, np.int32), row_time = np.zeros(0, np.float32),
                         latents = [], instances = [], modality_positions = [[] for _ in range(S)], total_tokens = S, n_type_tokens = [])
        rb.kv_row = (base + ln).astype(np.int32)
        res = self.eng.forward(rb, None, None, train = False, want_logits = True, cache = self.cache)
        V = self.eng.model.to_text_logits.weight.shape[0]
        self._update([res['logits'][s, :V].float() for s in range(S)], advance = 1)

    def run(self, max_steps):
        n = 0
        while n < max_steps:
            self.step(); n += 1
            if self.text_left == 0:
                break
        return n


def _oracle_new_cache(self, n_slabs, cap):
    return OracleKVCache(self.model, n_slabs, cap)


def _oracle_text_decoder(self, cache, S, **kw):
    return OracleTextDecoder(self, cache, S, **kw)


def _oracle_ode_solve(self, cache, rb, y, *, dup, steps, cfg_scale, use_graph = True):
    """fixed-grid midpoint (the torchdiffeq restatement of oracle/shims) with the cond | uncond batch layout of decode.ode_solve"""
    grid = torch.linspace(0, 1, steps)
    types = [t for t, v in enumerate(y) if v is not None]
    def flow(tval, ys):
        rb.cond_times[:] = float(tval)
        x = [torch.cat([v] * dup) if v is not None else None for v in ys]
        res = self.forward(rb, x, None, train = False, want_preds = True, cache = cache)
        out = []
        for t, v in enumerate(ys):
            if v is None:
                out.append(None); continue
            p = res['preds'][t]
            pc = p[:v.shape[0]]
            out.append(p[v.shape[0]:] + cfg_scale * (pc - p[v.shape[0]:]) if dup == 2 else pc)
        return out
    for t0, t1 in zip(grid[:-1], grid[1:]):
        dt = t1 - t0
        f0 = flow(t0, y)
        ymid = [v + f * (0.5 * dt) if v is not None else None for v, f in zip(y, f0)]
        f1 = flow(t0 + 0.5 * dt, ymid)
        y = [v + dt * f if v is not None else None for v, f in zip(y, f1)]
    return y


OracleEngine.new_cache = _oracle_new_cache
OracleEngine.text_decoder = _oracle_text_decoder
OracleEngine.ode_solve = _oracle_ode_solve
