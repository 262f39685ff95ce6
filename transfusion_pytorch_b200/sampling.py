"""Batched sampling over the B200 engine: `sample_many`, `sample_one` (= `sample`), `generate_modality_only`.

State machine of the reference's `sample_many` (transfusion.py:2079-2583): every sample alternates between a text
phase (one token per step, all text-phase samples share one forward) and a modality phase (one joint fixed-grid
midpoint ODE - torchdiffeq semantics, transfusion.py:1314-1318 - for all modality-phase samples, classifier-free
guidance `uncond + s (cond - uncond)`, transfusion.py:2521).

B200 design difference: there is no padded kv cache that is re-padded and concatenated on every step
(transfusion.py:2257-2277, 2323-2327, 2531-2533).  The sequences here are short (<= a few thousand tokens per
sample) and the ragged engine runs the whole packed batch in one pass, so every step recomputes the packed prefix.
To stay numerically IDENTICAL to the reference's cached path the recomputation reproduces what the cache held:
  * a prompt modality contributes keys/values of its latents at t = 1 (prefill, transfusion.py:2187-2201);
  * a modality decoded in this call contributes the keys/values of the LAST ODE evaluation - the midpoint state
    y_mid at time t_{N-2} + h/2 - because that is the cache the reference commits (transfusion.py:2466-2533);
  * the unconditional branch is rebuilt from the final samples at t = 1 (transfusion.py:2389-2406).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
from torch import tensor, cat, is_tensor

from .modality_processing import pack_batch, is_int_tensor


@dataclass
class _State:
    sample: list                                 # parts assembled so far: text tensors and (type, latents) tuples
    kv_src: list                                 # same structure; modalities carry (type, latents_for_kv, time)
    curr_seq: torch.Tensor
    phase: str = 'text'
    num_past_modalities: int = 0
    curr_modality_id: int | None = None
    modality_shape: tuple | None = None
    modality_length: int | None = None
    dim_latent: int | None = None
    num_tokens: int = 0


def _tokens_since_rightmost(t, rid):
    idx = (t == rid).nonzero()
    return t[0:0] if idx.numel() == 0 else t[int(idx[-1]) + 1:]


class SamplingMixin:
    # ------------------------------------------------------------------ prompt handling (transfusion.py:1701-1825)
    def _shape_from_seq(self, seq, modality_id, fixed_shape, forced_shape):
        shape = forced_shape if forced_shape is not None else fixed_shape
        meta = _tokens_since_rightmost(seq, self.meta_id)
        default_shape = self.modality_default_shape[modality_id]
        ndim = self.modality_num_dim[modality_id]
        if meta.numel() > 0 and forced_shape is None:
            meta_str = self.decode_chars(meta[:-1])
            if not meta_str.isdigit() or int(meta_str) <= 0:
                assert default_shape is not None, 'invalid modality meta information detected, please set `modality_default_shape` in order to properly fallback'
                shape = default_shape
            else:
                shape = self.to_modality_shape_fn[modality_id](meta_str)
        shape = shape if shape is not None else default_shape
        if self.fallback_to_default_shape_if_invalid and ndim is not None and len(shape) != ndim:
            shape = default_shape
        assert shape is not None, f'language model did not produce a proper modality shape for modality type {modality_id} - please set a fallback shape with `modality_default_shape`'
        assert ndim is None or ndim == len(shape), f'expected modality type {modality_id} to have {ndim} dimensions but language model produced a shape of {shape}'
        return tuple(shape)

    def _transition(self, seq, fixed_shape, forced_id, forced_shape):
        last = int(seq[-1])
        if last not in self.som_ids:
            return None
        mid = self.som_ids.index(last)
        return mid, self._shape_from_seq(seq, mid, fixed_shape, forced_shape if mid == forced_id else None)

    def prepare_prompt_sample(self, prompt, force_modality_at_start):
        if is_tensor(prompt) and prompt.is_floating_point():
            prompt = (0, prompt)
        if is_int_tensor(prompt):
            prompt = [prompt]
        elif isinstance(prompt, tuple):
            mtype, modality = prompt
            enc = self.modality_encoder[mtype]
            if enc is not None:
                with torch.no_grad():
                    enc.eval()
                    modality = (enc(modality[None])[0] if self.encdec_needs_batch_dim else enc(modality)).detach()
            cf = self.channel_first_latent[mtype]
            axial = tuple(modality.shape[1:]) if cf else tuple(modality.shape[:-1])
            prompt = [tensor([self.meta_id]), self.char_tokenizer(','.join(map(str, axial))), tensor([self.som_ids[mtype]]), (mtype, modality), tensor([self.eom_ids[mtype]])]
        elif prompt is None:
            prompt = []
        prompt = [p for p in prompt if p is not None]
        if prompt and not is_tensor(prompt[-1]):
            prompt.append(tensor([self.eom_ids[prompt[-1][0]]]))
        parts = [tensor([self.sos_id])]
        for p in prompt:                               # concat contiguous text
            if is_tensor(p) and not p.is_floating_point():
                p = p.reshape(-1).long().cpu()
                if is_tensor(parts[-1]):
                    parts[-1] = cat((parts[-1], p))
                else:
                    parts.append(p)
            else:
                p = (0, p) if is_tensor(p) else p
                parts.append((p[0], p[1].detach().float().cpu()))
        forced_id, forced_shape = force_modality_at_start if isinstance(force_modality_at_start, tuple) else (force_modality_at_start, None)
        if forced_id is not None:
            if forced_shape is not None:
                forced = cat((tensor([self.meta_id]), self.char_tokenizer(','.join(map(str, forced_shape))), tensor([self.som_ids[forced_id]])))
            else:
                forced = tensor([self.som_ids[forced_id]])
            if is_tensor(parts[-1]):
                parts[-1] = cat((parts[-1], forced))
            else:
                parts.append(forced)
        return parts, forced_id, forced_shape

    # ------------------------------------------------------------------ engine passes
    def _decode_forward(self, seqs, want_logits):
        """seqs: list of part lists whose modalities are (type, latents [L, dl] or shaped, time).  One ragged forward."""
        samples, times = [], []
        for parts in seqs:
            s, ts = [], []
            for p in parts:
                if is_tensor(p):
                    s.append(p)
                else:
                    s.append((p[0], p[1])); ts.append(float(p[2]))
            samples.append(s); times.append(ts)
        m = max((len(t) for t in times), default = 0)
        tm = torch.zeros(len(seqs), max(m, 1))
        for i, t in enumerate(times):
            if t:
                tm[i, :len(t)] = tensor(t)
        rb = pack_batch(samples, tm if m > 0 else None, self, return_loss = False, return_embed = True)
        lat = self._latents_to_device(rb)
        res = self.engine.forward(rb, lat, None, train = False, want_logits = True)
        return rb, res

    def _last_instance_rows(self, rb):
        """compact row range of the LAST modality instance of every sample"""
        last = {}
        for inst in rb.instances:
            last[inst.batch_index] = inst
        out = []
        for b in range(rb.B):
            inst = last[b]
            r0 = rb.type_rows[inst.modality_type][0] + inst.row0
            out.append((inst.modality_type, r0, r0 + inst.length))
        return out

    # ------------------------------------------------------------------ sample_many (transfusion.py:2079-2583)
    @torch.no_grad()
    def sample_many(self, prompts = None, max_length = 2048, text_temperature = 1.0, text_min_p = 0.1, fixed_modality_shape = None,
                    force_modality_at_start = None, init_modality_noise = None, modality_steps = 16, return_unprocessed_modalities = False, cfg_scale = 3.):
        from .transfusion import sample_text_token
        was_training = self.training
        self.eval()
        try:
            if prompts is None:
                prompts = [None]
            elif not isinstance(prompts, list):
                prompts = [prompts]
            states, forced_id, forced_shape = [], None, None
            for prompt in prompts:
                parts, forced_id, forced_shape = self.prepare_prompt_sample(prompt, force_modality_at_start)
                kv = [p if is_tensor(p) else (p[0], p[1], 1.0) for p in parts]          # prompt modalities are conditioned at t = 1
                last = parts[-1]
                st = _State(sample = parts, kv_src = kv, curr_seq = last if is_tensor(last) else tensor([self.sos_id]))
                st.num_past_modalities = sum(not is_tensor(p) for p in parts)
                states.append(st)

            def maybe_transition(st):
                tr = self._transition(st.curr_seq, fixed_modality_shape, forced_id, forced_shape)
                if tr is None:
                    return False
                st.curr_modality_id, st.modality_shape = tr
                st.modality_length = math.prod(st.modality_shape)
                st.dim_latent = self.dim_latents[st.curr_modality_id]
                st.phase = 'modality'
                return True

            for st in states:
                maybe_transition(st)

            def step_text(group):
                rb, res = self._decode_forward([st.kv_src for st in group], True)
                V = self.to_text_logits.weight.shape[0]
                last_rows = torch.as_tensor(rb.cu[1:] - 1, device = res['logits'].device)
                logits = res['logits'][last_rows, :V].float()
                sampled = sample_text_token(logits, text_temperature, text_min_p).cpu()
                for st, tok in zip(group, sampled):
                    st.curr_seq = cat((st.curr_seq, tok))
                    st.sample[-1] = st.curr_seq
                    st.kv_src[-1] = st.curr_seq
                    st.num_tokens += 1
                    if int(tok) == self.eos_id:
                        st.phase = 'done'; continue
                    if st.num_tokens > max_length:
                        st.phase = 'done'; continue
                    maybe_transition(st)

            def step_modality(group):
                dev = self.device
                use_cfg = cfg_scale != 1.
                ys = []
                for st in group:
                    L, dl = st.modality_length, st.dim_latent
                    noise = init_modality_noise[:L, :dl] if init_modality_noise is not None else torch.randn(L, dl)
                    assert noise.shape == (L, dl)
                    ys.append(noise.float().to(dev))
                uncond_hist = [[torch.full_like(p, self.null_text_id) if is_tensor(p) else (p[0], p[1], 1.0) for p in st.sample] for st in group]
                last_eval = {}

                def flows(t, ys_now, record):
                    tval = float(t)
                    cond_seqs = [[*st.kv_src, (st.curr_modality_id, y, tval)] for st, y in zip(group, ys_now)]
                    rb, res = self._decode_forward(cond_seqs, True)
                    rows = self._last_instance_rows(rb)
                    cond = [res['preds'][ty][r0 - rb.type_rows[ty][0]: r1 - rb.type_rows[ty][0]].clone() for ty, r0, r1 in rows]
                    if record:
                        last_eval['y'], last_eval['t'] = [y.clone() for y in ys_now], tval
                    if not use_cfg:
                        return cond
                    unc_seqs = [[*h, (st.curr_modality_id, y, tval)] for st, h, y in zip(group, uncond_hist, ys_now)]
                    rb2, res2 = self._decode_forward(unc_seqs, True)
                    rows2 = self._last_instance_rows(rb2)
                    unc = [res2['preds'][ty][r0 - rb2.type_rows[ty][0]: r1 - rb2.type_rows[ty][0]] for ty, r0, r1 in rows2]
                    return [u + cfg_scale * (c - u) for c, u in zip(cond, unc)]

                grid = torch.linspace(0, 1, modality_steps).tolist()
                for k in range(len(grid) - 1):                     # fixed-grid explicit midpoint (torchdiffeq 'midpoint')
                    t0, h = grid[k], grid[k + 1] - grid[k]
                    f0 = flows(t0, ys, False)
                    y_mid = [y + f * (0.5 * h) for y, f in zip(ys, f0)]
                    f1 = flows(t0 + 0.5 * h, y_mid, True)
                    ys = [y + f * h for y, f in zip(ys, f1)]
                for st, y, ymid in zip(group, ys, last_eval.get('y', ys)):
                    cf = self.channel_first_latent[st.curr_modality_id]
                    shaped = y.reshape(*st.modality_shape, st.dim_latent)
                    if cf:
                        shaped = shaped.movedim(-1, 0)
                    st.sample.append((st.curr_modality_id, shaped))
                    kv_lat = ymid.reshape(*st.modality_shape, st.dim_latent)
                    st.kv_src.append((st.curr_modality_id, kv_lat.movedim(-1, 0) if cf else kv_lat, last_eval.get('t', 1.0)))
                    st.curr_seq = tensor([self.eom_ids[st.curr_modality_id]])
                    st.sample.append(st.curr_seq); st.kv_src.append(st.curr_seq)
                    st.num_tokens += st.modality_length
                    st.num_past_modalities += 1
                    st.phase = 'text'
                    if st.num_tokens > max_length:
                        st.phase = 'done'

            while not all(st.phase == 'done' for st in states):
                text_group = [st for st in states if st.phase == 'text']
                while text_group:
                    step_text(text_group)
                    text_group = [st for st in text_group if st.phase == 'text']
                mod_group = [st for st in states if st.phase == 'modality']
                while mod_group:
                    step_modality(mod_group)
                    mod_group = [st for st in mod_group if st.phase == 'modality']

            samples = [st.sample for st in states]
            if return_unprocessed_modalities:
                return samples
            return self.decode_modalities(samples)
        finally:
            self.train(was_training)

    def decode_modalities(self, samples):
        out = []
        for s in samples:
            parts = []
            for p in s:
                if not is_tensor(p) and self.modality_decoder[p[0]] is not None:
                    dec = self.modality_decoder[p[0]]
                    dec.eval()
                    v = p[1].to(self.device)
                    p = (p[0], dec(v[None])[0] if self.encdec_needs_batch_dim else dec(v))
                parts.append(p)
            out.append(parts)
        return out

    @torch.no_grad()
    def sample_one(self, prompt = None, max_length = 2048, text_temperature = 1.0, text_min_p = 0.1, cache_kv = False, fixed_modality_shape = None,
                   force_modality_at_start = None, init_modality_noise = None, modality_steps = 16, return_unprocessed_modalities = False, cfg_scale = 3.):
        """single-sample sampling; follows the `cache_kv = True` semantics of the reference (the path `sample_many` mirrors)."""
        if self.num_text_tokens == 0:
            fid, fshape = force_modality_at_start if isinstance(force_modality_at_start, tuple) else (force_modality_at_start, None)
            return self.generate_modality_only(batch_size = 1, modality_type = fid, fixed_modality_shape = fshape)
        return self.sample_many([prompt], max_length = max_length, text_temperature = text_temperature, text_min_p = text_min_p,
                                fixed_modality_shape = fixed_modality_shape, force_modality_at_start = force_modality_at_start,
                                init_modality_noise = init_modality_noise, modality_steps = modality_steps,
                                return_unprocessed_modalities = return_unprocessed_modalities, cfg_scale = cfg_scale)[0]

    sample = sample_one

    @torch.no_grad()
    def generate_modality_only(self, batch_size = 1, modality_type = None, fixed_modality_shape = None, modality_steps = 16, return_unprocessed_modalities = False,
                               init_noise = None):
        """transfusion.py:2868-2923: midpoint ODE over `forward_modality(..., return_loss = False)`"""
        was_training = self.training
        self.eval()
        try:
            if self.num_modalities > 1:
                assert modality_type is not None, '`modality_type` must be explicitly passed in on forward when training on greater than 1 modality'
            mt = modality_type if modality_type is not None else 0
            shape = fixed_modality_shape if fixed_modality_shape is not None else self.modality_default_shape[mt]
            assert shape is not None
            dl = self.dim_latents[mt]
            y = init_noise.float() if init_noise is not None else torch.randn(batch_size, *shape, dl)
            if self.channel_first_latent[mt]:
                y = y.movedim(-1, 1)
            y = y.to(self.device)
            grid = torch.linspace(0., 1., modality_steps).tolist()
            f = lambda t, v: self.forward_modality(v, times = torch.full((batch_size,), t), modality_type = mt, encode_modality = False, return_loss = False)
            for k in range(len(grid) - 1):
                t0, h = grid[k], grid[k + 1] - grid[k]
                y_mid = y + f(t0, y) * (0.5 * h)
                y = y + f(t0 + 0.5 * h, y_mid) * h
            dec = self.modality_decoder[mt]
            if dec is not None and not return_unprocessed_modalities:
                dec.eval(); y = dec(y)
            return y
        finally:
            self.train(was_training)
