"""Batched sampling over the B200 engine: `sample_many`, `sample_one` (= `sample`), `generate_modality_only`.

State machine of the reference's `sample_many` (transfusion.py:2079-2583): every sample alternates between a text
phase (one token per step, all text-phase samples share one forward) and a modality phase (one joint fixed-grid
midpoint ODE - torchdiffeq semantics, transfusion.py:1314-1318 - for all modality-phase samples, classifier-free
guidance `uncond + s (cond - uncond)`, transfusion.py:2521).

B200 design: the kv cache is a set of in-place slabs (`engine.KVCache`) instead of per-sample tensors that are re-padded and
concatenated on every step (transfusion.py:2257-2277, 2323-2327, 2531-2533); the text loop and the ODE loop run as captured CUDA
graphs over device-resident sampler state (`decode.py`, csrc/decode.cu); the conditional and unconditional branch of an ODE
evaluation share one ragged forward.  What the cache HOLDS is exactly what the reference's holds:
  * a prompt modality contributes keys/values of its latents at t = 1 (prefill, transfusion.py:2187-2201);
  * a modality decoded in this call leaves the keys/values of the LAST ODE evaluation - the midpoint state at t_{N-2} + h/2 -
    because that is the cache the reference commits (transfusion.py:2466-2533);
  * a sampled [som] token never gets a cache row: the modality takes its RoPE position (transfusion.py:2337-2349, 2408-2411);
  * the unconditional branch is rebuilt per modality round from the null-id history with earlier modalities at t = 1
    (transfusion.py:2385-2406).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
from torch import tensor, cat, is_tensor

import numpy as np

from .modality_processing import pack_batch, pack_incremental, is_int_tensor


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


@dataclass(eq = False)
class _State:
    sample: list                                 # parts assembled so far: text tensors and (type, latents) tuples
    curr_seq: torch.Tensor                       # the running text part (tokens since the last modality)
    phase: str = 'text'
    last_token: int = 0                          # fed to the next text step (it gets its cache row then, T.py:2246-2248)
    cache_len: int = 0                           # rows of this sample's cache slab that are committed
    tokens_seen: int = 0                         # RoPE position of the next token (a whole modality counts as one, T.py:2332, 2548)
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

    def _modality_len(self, part):
        cf = self.channel_first_latent[part[0]]
        return math.prod(part[1].shape[1:] if cf else part[1].shape[:-1])

    def _parts_len(self, parts):
        return sum(p.numel() if is_tensor(p) else self._modality_len(p) for p in parts)

    # ------------------------------------------------------------------ sample_many (transfusion.py:2079-2583)
    @torch.no_grad()
    def sample_many(self, prompts = None, max_length = 2048, text_temperature = 1.0, text_min_p = 0.1, fixed_modality_shape = None,
                    force_modality_at_start = None, init_modality_noise = None, modality_steps = 16, return_unprocessed_modalities = False, cfg_scale = 3.,
                    seed = None, use_cuda_graph = True):
        """Batched sampling with a real kv cache.  Follows the reference's `sample_many` state machine step for step - prefill of all prompts in
        one ragged forward, a shared text loop, a joint midpoint ODE per modality round with classifier-free guidance - but the cache is a set
        of in-place slabs, the text loop and the ODE loop are captured CUDA graphs over device-resident state, and conditional + unconditional
        branches of an ODE evaluation share one forward.  `seed` (extension) seeds the on-device token sampler (default: drawn from torch's
        global generator); `use_cuda_graph = False` launches the same kernels eagerly."""
        from .decode import ST_LEN, ST_SEEN, ST_LAST, ST_PHASE, ST_NTOK, PH_TEXT, PH_MODALITY, PH_DONE
        was_training = self.training
        self.eval()
        eng = self.engine
        frozen_before = getattr(eng, 'frozen', False)
        try:
            if prompts is None:
                prompts = [None]
            elif not isinstance(prompts, list):
                prompts = [prompts]
            states, forced_id, forced_shape = [], None, None
            for prompt in prompts:
                parts, forced_id, forced_shape = self.prepare_prompt_sample(prompt, force_modality_at_start)
                last = parts[-1]
                st = _State(sample = parts, curr_seq = last if is_tensor(last) else tensor([self.sos_id]))
                st.last_token = int(last[-1]) if is_tensor(last) else 0
                st.num_past_modalities = sum(not is_tensor(p) for p in parts)
                # tokens in the cache after the prefill and the RoPE position of the next token: every modality collapses to ONE position
                seq_len = self._parts_len(parts)
                st.cache_len, st.tokens_seen = seq_len, seq_len - sum(self._modality_len(p) - 1 for p in parts if not is_tensor(p))
                states.append(st)
            S = len(states)
            use_cfg = cfg_scale != 1.

            def maybe_transition(st):
                tr = self._transition(st.curr_seq, fixed_modality_shape, forced_id, forced_shape)
                if tr is None:
                    return False
                st.curr_modality_id, st.modality_shape = tr
                st.modality_length = math.prod(st.modality_shape)
                st.dim_latent = self.dim_latents[st.curr_modality_id]
                st.phase = 'modality'
                return True

            # ---- slab capacity: prompt + generated tokens + the largest modality that can be decoded (grown on demand if a sampled shape is larger)
            shapes = [s_ for s_ in self.modality_default_shape if s_ is not None] + [s_ for s_ in (fixed_modality_shape, forced_shape) if s_ is not None]
            mod_guess = max([math.prod(s_) for s_ in shapes], default = 0)
            cap = _round_up(max(st.cache_len for st in states) + max_length + mod_guess + 8, 64)
            cache = eng.new_cache(2 * S if use_cfg else S, cap)
            eng.pack_weights()
            eng.frozen = True                                  # parameters cannot change inside a no-grad sampling call

            def ensure_cap(need):
                nonlocal cache, cap, dec
                if need <= cap:
                    return
                new_cap = _round_up(need + 64, 64)
                bigger = eng.new_cache(cache.n_slabs, new_cap)
                for name in ('k', 'v'):
                    src, dst = getattr(cache, name), getattr(bigger, name)
                    dst.view(dst.shape[0], cache.n_slabs, new_cap, -1)[:, :, :cap].copy_(src.view(src.shape[0], cache.n_slabs, cap, -1))
                cache, cap = bigger, new_cap
                dec = None                                     # the decoder (and its captured graph) is bound to the old slabs

            # ---- prefill: ONE ragged forward over all prompts builds every sample's cache and the logits of its first token (T.py:2176-2201)
            m = max((st.num_past_modalities for st in states), default = 0)
            times = torch.ones(S, m) if m > 0 else None        # prompted modalities are conditioned at t = 1 (T.py:2187-2192)
            rb = pack_incremental([st.sample for st in states], times, self, slab = np.arange(S), base_len = np.zeros(S), rope_base = np.zeros(S), cap = cap)
            res = eng.forward(rb, self._latents_to_device(rb), None, train = False, want_logits = True, cache = cache)
            for st in states:
                maybe_transition(st)                           # a prompt ending in [som] starts with the modality phase

            if seed is None:
                seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
            dec = None
            def decoder():
                nonlocal dec
                if dec is None:
                    dec = eng.text_decoder(cache, S, slab0 = 0, hist_cap = max_length + 4, eos_id = self.eos_id, som_ids = self.som_ids, max_length = max_length,
                                           temperature = text_temperature, min_p = text_min_p, vlimit = 0, seed = seed, use_graph = use_cuda_graph)
                return dec

            def push_state():
                ph = dict(text = PH_TEXT, modality = PH_MODALITY, done = PH_DONE)
                decoder().set_state([st.cache_len for st in states], [st.tokens_seen for st in states], [st.last_token for st in states],
                                    [ph[st.phase] for st in states], [st.num_tokens for st in states])

            def pull_state():
                """device sampler state -> host `_State`s (tokens sampled since the last push are appended to the running text)"""
                arr, hist = decoder().get_state()
                for i, st in enumerate(states):
                    if st.phase != 'text':
                        continue
                    new = torch.from_numpy(hist[i])
                    if new.numel():
                        st.curr_seq = cat((st.curr_seq, new))
                        st.sample[-1] = st.curr_seq
                        st.last_token = int(new[-1])
                    st.cache_len, st.tokens_seen, st.num_tokens = int(arr[ST_LEN, i]), int(arr[ST_SEEN, i]), int(arr[ST_NTOK, i])
                    p = int(arr[ST_PHASE, i])
                    if p == PH_DONE:
                        st.phase = 'done'
                    elif p == PH_MODALITY:
                        assert maybe_transition(st)

            # first token of every text-phase sample, from the prefill logits at its last prompt position (T.py:2225-2250)
            push_state()
            decoder().sample_first(res['logits'], rb.cu[1:] - 1)
            pull_state()

            def text_loop():
                """all samples in the text phase advance in lock-step until each has hit [eos], the length limit or a [som] (T.py:2563-2568)"""
                push_state()
                budget = max(1, max_length + 2 - min(st.num_tokens for st in states if st.phase == 'text'))
                decoder().run(budget)
                pull_state()

            def step_modality(group):
                dev = self.device
                G = len(group)
                ensure_cap(max(st.cache_len + st.modality_length for st in group) + 1)
                # initial noise per sample, gathered per modality type in scan order (the engine's compact row order)
                ys = []
                for st in group:
                    L, dl = st.modality_length, st.dim_latent
                    noise = init_modality_noise[:L, :dl] if init_modality_noise is not None else torch.randn(L, dl)
                    assert noise.shape == (L, dl)
                    ys.append(noise.float())
                slabs = [states.index(st) for st in group]
                base, ropes = [st.cache_len for st in group], [st.tokens_seen for st in group]
                # (the true axial shape, not the flattened length: the axial positional embedding reads the coordinates off it)
                placeholder = lambda st: torch.empty(st.dim_latent, *st.modality_shape) if self.channel_first_latent[st.curr_modality_id] else torch.empty(*st.modality_shape, st.dim_latent)
                samples = [[(st.curr_modality_id, placeholder(st))] for st in group]      # shapes only: the latents live on the device (ODE state)
                if use_cfg:
                    # unconditional branch: every text token replaced by the null id, earlier modalities at t = 1 - rebuilt per modality round
                    # into the second half of the slabs by one ragged prefill over the whole group (T.py:2385-2406)
                    hist = [[torch.full_like(p, self.null_text_id) if is_tensor(p) else p for p in st.sample] for st in group]
                    ulen = [self._parts_len(h) for h in hist]
                    ensure_cap(max(u + st.modality_length for u, st in zip(ulen, group)) + 1)
                    mu = max(st.num_past_modalities for st in group)
                    urb = pack_incremental(hist, torch.ones(G, mu) if mu > 0 else None, self, slab = np.asarray(slabs) + S, base_len = np.zeros(G), rope_base = np.zeros(G), cap = cap)
                    eng.forward(urb, self._latents_to_device(urb), None, train = False, want_logits = False, want_preds = False, cache = cache)
                    slabs = slabs + [s_ + S for s_ in slabs]
                    base, ropes, samples = base + ulen, ropes + ropes, samples + samples
                dup = 2 if use_cfg else 1
                orb = pack_incremental(samples, torch.zeros(len(samples), 1), self, slab = np.asarray(slabs), base_len = np.asarray(base), rope_base = np.asarray(ropes), cap = cap)
                # per type: conditional rows of the group in scan order; the unconditional copies follow (pack order = sample order)
                y = []
                for t in range(self.num_modalities):
                    rows = [ys[i] for i, st in enumerate(group) if st.curr_modality_id == t]
                    y.append(cat(rows).to(dev).contiguous() if rows else None)
                y = eng.ode_solve(cache, orb, y, dup = dup, steps = modality_steps, cfg_scale = cfg_scale, use_graph = use_cuda_graph)
                # commit: the cache rows written by the LAST evaluation stay (T.py:2529-2533), the final state is the sampled modality
                offs = [0] * self.num_modalities
                finals = [v.cpu() if v is not None else None for v in y]
                for st in group:
                    t, L = st.curr_modality_id, st.modality_length
                    sampled = finals[t][offs[t]: offs[t] + L]
                    offs[t] += L
                    shaped = sampled.reshape(*st.modality_shape, st.dim_latent)
                    if self.channel_first_latent[t]:
                        shaped = shaped.movedim(-1, 0)
                    st.sample.append((t, shaped))
                    st.curr_seq = tensor([self.eom_ids[t]])
                    st.sample.append(st.curr_seq)
                    st.last_token = self.eom_ids[t]
                    st.cache_len += L
                    st.tokens_seen += 1
                    st.num_tokens += L
                    st.num_past_modalities += 1
                    st.phase = 'done' if st.num_tokens > max_length else 'text'

            # optional phase timing (bench.py: `model._sampling_timer = {}`): CUDA events around every text loop / modality round
            timer = getattr(self, '_sampling_timer', None)
            def timed(name, fn, *a):
                if timer is None:
                    return fn(*a)
                e0, e1 = torch.cuda.Event(enable_timing = True), torch.cuda.Event(enable_timing = True)
                e0.record(); fn(*a); e1.record()
                timer.setdefault(name, []).append((e0, e1))
            while not all(st.phase == 'done' for st in states):
                if any(st.phase == 'text' for st in states):
                    timed('text_loop', text_loop)
                group = [st for st in states if st.phase == 'modality']
                if group:
                    timed('modality_round', step_modality, group)

            samples = [st.sample for st in states]
            if return_unprocessed_modalities:
                return samples
            return self.decode_modalities(samples)
        finally:
            eng.frozen = frozen_before
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
