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

    def stack(self, rb, x0, cond_tok, is_mod, kv_limit, rope_pos, valid):
        """x0 [b,n,d]; cond_tok [b,n] time per token or None; returns final-norm output and hiddens."""
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
            sim = torch.einsum('bhid,bhjd->bhij', q * 64 ** -0.5, k)
            cap = tr.softcap_value
            sim = (sim / cap).tanh() * cap                                                  # T.py:1001
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

    def run(self, rb, latents, eps, *, text_loss_weight = 1., flow_loss_weight = 1., vlimit = 0, modality_only = False, want_loss = True):
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
        out, hid = self.stack(rb, x0, cond_tok, is_mod, kv_limit, rope_pos, valid)
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
        self.ref = TorchReference(model)
        self.model = model
        self.state = None

    def forward(self, rb, latents, eps, *, train, want_logits = False, vlimit = 0, text_loss_weight = 1., flow_loss_weight = 1., modality_only = False):
        with torch.set_grad_enabled(train):
            res = self.ref.run(rb, latents, eps, text_loss_weight = text_loss_weight, flow_loss_weight = flow_loss_weight, vlimit = vlimit,
                               modality_only = modality_only, want_loss = train)
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
