"""Device-resident decode loops over the slab kv cache (`engine.KVCache`): the text loop and the flow-matching ODE loop of
`sample_many` / `generate_text_only` (reference transfusion.py:2279-2349 `step_text`, 2354-2556 `step_modality`, 2669-2707).

Both loops are ONE captured CUDA graph replayed from device-resident state:

  text step   tfx_decode_prep (sampler state -> token metadata) -> block stack over S new tokens with in-place kv append and the
              single-query decode attention -> logits GEMM -> tfx_sample_tokens (argmax / min-p + Gumbel, state machine update).
              The host replays the graph `poll` times between reads of ONE int (samples still in the text phase): no per-token `.item()`.
  ODE eval    tfx_ode_pre (y + c f_prev into the model input, step time into the condition table) -> block stack over the modality
              tokens of the conditional AND unconditional branch in one ragged batch (each branch attends its own cache slabs) ->
              flow head -> tfx_ode_post (classifier-free-guidance combine + midpoint update).  Replayed 2 (steps - 1) times.

The host-side state machine that decides WHEN these loops run lives in `sampling.py`; the CPU test double with the same interface is
`oracle/torch_reference.OracleEngine` (tests only).
"""
from __future__ import annotations

import numpy as np
import torch

from .modality_processing import RaggedBatch

I32, F32 = torch.int32, torch.float32

# sampler state rows (include/tfx_b200.h)
ST_LEN, ST_SEEN, ST_LAST, ST_PHASE, ST_NTOK, ST_HIST = range(6)
PH_TEXT, PH_MODALITY, PH_DONE = 0, 1, 2


class TextDecoder:
    """S samples decoding text in lock-step against cache slabs [slab0, slab0 + S)."""

    def __init__(self, engine, cache, S, *, slab0 = 0, hist_cap, eos_id, som_ids, max_length, temperature, min_p, vlimit = 0, seed = 0, use_graph = True, poll = 8):
        self.eng, self.cache, self.S, self.slab0 = engine, cache, int(S), int(slab0)
        dev = engine.device
        self.hist_cap = int(hist_cap)
        self.state = torch.zeros(6, S, device = dev, dtype = I32)
        self.hist = torch.zeros(S, self.hist_cap, device = dev, dtype = I32)
        self.counters = torch.zeros(2, device = dev, dtype = I32)
        self.som = torch.tensor(list(som_ids) if som_ids else [-1], device = dev, dtype = I32)
        self.n_som = len(som_ids) if som_ids else 0
        self.eos_id, self.max_length = int(eos_id), int(max_length)
        self.temperature, self.min_p, self.vlimit, self.seed = float(temperature), float(min_p), int(vlimit), int(seed) & (2 ** 63 - 1)
        self.use_graph, self.poll = use_graph, max(1, int(poll))
        self.graph = None
        self.launches_per_step = 0
        # static descriptor of one text step: S new tokens, one per sample; all metadata is written by tfx_decode_prep
        self.meta = torch.zeros(8, S, device = dev, dtype = I32)
        z = np.zeros(S, dtype = np.int32)
        rb = RaggedBatch(B = S, M = S, seq_lens = np.ones(S, dtype = np.int64), cu = np.arange(S + 1, dtype = np.int64), full_lens = np.ones(S, dtype = np.int64),
                         text_id = z, label = z, kv_limit = z, rope_pos = z, cond_row = z, slot = z, n_cond = 0, cond_times = np.zeros(0, np.float32), n_types = 0,
                         type_rows = [], row_token = np.zeros(0, np.int32), row_time = np.zeros(0, np.float32), latents = [], instances = [],
                         modality_positions = [[] for _ in range(S)], total_tokens = S, n_type_tokens = [])
        for n in ('tile_q0', 'tile_qend', 'tile_kv0', 'tile_kvend', 't2_q0', 't2_qend', 't2_kv0', 't2_kvend'):
            setattr(rb, n, z)
        rb.single_row_tiles = True
        rb.max_rope_pos = cache.cap + 1
        m = self.meta
        rb.dev = dict(text_id = m[0], rope_pos = m[1], kv_row = m[2], kv_limit = m[3], tile_q0 = m[4], tile_qend = m[5], tile_kv0 = m[6], tile_kvend = m[7],
                      t2_q0 = m[4], t2_qend = m[5], t2_kv0 = m[6], t2_kvend = m[7], h2d_bytes = 0)
        self.rb = rb

    # ---- state transfer (host <-> device), a handful of ints per sample
    def set_state(self, length, tokens_seen, last_token, phase, num_tokens):
        host = torch.zeros(6, self.S, dtype = I32)
        for r, a in ((ST_LEN, length), (ST_SEEN, tokens_seen), (ST_LAST, last_token), (ST_PHASE, phase), (ST_NTOK, num_tokens)):
            host[r] = torch.as_tensor(np.asarray(a, dtype = np.int64)).to(I32)
        self.state.copy_(host.pin_memory(), non_blocking = True)
        self.hist.zero_()

    def update_rows(self, rows: dict):
        """overwrite whole state rows (e.g. after a modality phase): {ST_*: int array [S]}; the history cursor is reset"""
        host = self.state.cpu()
        for r, a in rows.items():
            host[r] = torch.as_tensor(np.asarray(a, dtype = np.int64)).to(I32)
        host[ST_HIST] = 0
        self.state.copy_(host.pin_memory(), non_blocking = True)

    def get_state(self):
        """(state int64 [6, S] on the host, list of the tokens sampled since the last reset per sample)"""
        st = self.state.cpu().long().numpy()
        hist = self.hist.cpu().numpy()
        assert (st[ST_HIST] <= self.hist_cap).all(), 'token history overflow'
        return st, [hist[s, :st[ST_HIST, s]].astype(np.int64) for s in range(self.S)]

    # ---- sampling of the first token from the prefill logits (T.py:2225-2250): the token is recorded but gets its cache row on the next step
    def sample_first(self, logits, rows):
        o, e = self.eng.ops, self.eng
        rows_dev = torch.as_tensor(np.asarray(rows, dtype = np.int32)).to(e.device)
        o.sample_tokens(logits, logits.shape[1], rows_dev, e.V, self.vlimit, self.state, self.S, self.hist, self.hist_cap, self.eos_id, self.som, self.n_som, self.max_length,
                        self.temperature, self.min_p, self.seed, self.counters, 0)

    def _step(self):
        o, e, m = self.eng.ops, self.eng, self.meta
        o.decode_prep(self.state, self.S, self.cache.cap, self.slab0, m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], self.counters)
        res = e.forward(self.rb, None, None, train = False, want_logits = True, cache = self.cache)
        lg = res['logits']
        o.sample_tokens(lg, lg.shape[1], None, e.V, self.vlimit, self.state, self.S, self.hist, self.hist_cap, self.eos_id, self.som, self.n_som, self.max_length,
                        self.temperature, self.min_p, self.seed, self.counters, 1)

    def step(self):
        """one text step for every sample in the text phase (captured after the first eager step of this decoder)"""
        e = self.eng
        if not self.use_graph:
            return self._step()
        if self.graph is None:
            l0 = e.ops.launches
            self._step()                                     # eager: allocates / sizes every workspace buffer of this shape
            self.launches_per_step = e.ops.launches - l0
            e.pin_workspaces()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            l0 = e.ops.launches
            with torch.cuda.graph(g):
                self._step()
            e.ops.launches = l0                              # the capture executed nothing
            self.graph = g
            self._pins = (dict(e.ws), dict(e.packed), e.fastp)
            return
        self.graph.replay()
        e.ops.launches += self.launches_per_step

    def run(self, max_steps):
        """text steps until no sample is left in the text phase (or max_steps); ONE 4-byte read per `poll` steps.  Returns the steps run."""
        n = 0
        while n < max_steps:
            k = min(self.poll, max_steps - n)
            for _ in range(k):
                self.step()
            n += k
            if int(self.counters[0].item()) == 0:
                break
        return n


def midpoint_table(steps: int, device):
    """(t, c, h, mode) per model evaluation of the fixed-grid midpoint solver on linspace(0, 1, steps), computed in fp32 exactly like
    torchdiffeq does (T.py:2523-2525): evaluation 2k at (t_k, y), evaluation 2k+1 at (t_k + dt/2, y + dt/2 f)."""
    grid = torch.linspace(0, 1, steps)
    rows = []
    for t0, t1 in zip(grid[:-1], grid[1:]):
        dt = t1 - t0
        rows.append([float(t0), 0.0, float(dt), 0.0])
        rows.append([float(t0 + 0.5 * dt), float(0.5 * dt), float(dt), 1.0])
    return torch.tensor(rows, dtype = F32).reshape(-1, 4).to(device)


def ode_solve(engine, cache, rb, y: list, *, dup: int, steps: int, cfg_scale: float, use_graph = True):
    """Integrates the flow ODE for the modality tokens described by `rb` (an uploaded incremental descriptor whose compact rows are, per
    type, `dup` copies of the group's tokens: conditional branch first, then - dup == 2 - the unconditional one).  `y[t]`: fp32 device
    matrix [n_t, dl_t] of initial noise for type t (None for absent types); integrated in place and returned."""
    o = engine.ops
    dev = engine.device
    n_evals = 2 * (steps - 1)
    if n_evals <= 0:
        return y
    tab = midpoint_table(steps, dev)
    idx = torch.zeros(1, device = dev, dtype = I32)
    types = [t for t, v in enumerate(y) if v is not None]
    x_eval = [torch.empty(dup * v.shape[0], v.shape[1], device = dev, dtype = F32) if v is not None else None for v in y]
    fprev = [torch.zeros_like(v) if v is not None else None for v in y]
    engine.upload(rb)
    dv = rb.dev

    def one_eval():
        for j, t in enumerate(types):
            n = y[t].numel()
            o.ode_pre(y[t], fprev[t], x_eval[t], n, dup, tab, idx, dv['cond_times'] if j == 0 else None, rb.n_cond if j == 0 else 0)
        res = engine.forward(rb, x_eval, None, train = False, want_logits = False, want_preds = True, cache = cache)
        for t in types:
            pred = res['preds'][t]
            n = y[t].numel()
            pc = pred[:y[t].shape[0]]
            pu = pred[y[t].shape[0]:] if dup == 2 else None
            o.ode_post(y[t], fprev[t], pc, pu, float(cfg_scale), n, tab, idx)
        o.counter_inc(idx)

    l0 = engine.ops.launches
    one_eval()
    per_eval = engine.ops.launches - l0
    if n_evals == 1:
        return y
    if not use_graph:
        for _ in range(n_evals - 1):
            one_eval()
        return y
    engine.pin_workspaces()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    l0 = engine.ops.launches
    with torch.cuda.graph(g):
        one_eval()
    engine.ops.launches = l0
    for _ in range(n_evals - 1):
        g.replay()
    engine.ops.launches += per_eval * (n_evals - 1)
    torch.cuda.current_stream().synchronize()                # the graph object (and its private pool) is dropped on return
    return y
