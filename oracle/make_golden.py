"""TEST INFRASTRUCTURE ONLY - generates tests/golden/*.pt by running the UNMODIFIED reference
(/root/reference, imported through oracle/reference_loader.py with the shims) on the deterministic
synthetic inputs of transfusion_pytorch_b200/synth.py.  Runs in the build container only
(`python -m oracle.make_golden`); the fixtures it writes are committed and are what the GPU tests read.

The reference has no golden vectors of its own: these fixtures ARE the pin of the
oracle - outputs of the reference itself on seeded inputs.
"""
from __future__ import annotations

import os
import sys
from unittest import mock

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.reference_loader import load_reference          # noqa: E402
from transfusion_pytorch_b200 import synth                   # noqa: E402

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def grad_fingerprint(model):
    """Per-parameter gradient fingerprint: [sum, abs-sum, projection on a fixed pseudo-random vector, l2] + first 8 values."""
    out = {}
    for name, p in model.named_parameters():
        if p.grad is None:
            continue
        g = p.grad.detach().float().reshape(-1)
        gen = torch.Generator().manual_seed(1234)
        proj = torch.randn(g.numel(), generator = gen)
        out[name] = dict(stats = torch.stack([g.sum(), g.abs().sum(), (g * proj).sum(), g.norm()]).double(), head = g[:8].clone())
    return out


def count_modalities(batch):
    return max(sum(1 for p in s if isinstance(p, tuple) or (torch.is_tensor(p) and p.is_floating_point())) for s in batch)


def noise_for(rb_like_rows, dl, seed):
    return torch.randn(rb_like_rows, dl, generator = torch.Generator().manual_seed(seed))


def run_interleaved(ref, name, ctor, batch, times, seed, subsample_rows = None, keep_hiddens = True):
    torch.manual_seed(0)
    model = ref.Transfusion(**ctor, modality_processing = 'flat')
    synth.fill_parameters_(model, seed = seed)
    model.eval()                                   # no CFG dropout; nothing else depends on the mode
    n_types = model.num_modalities
    # deterministic noise: the flat strategy draws ONE randn_like per modality type, on the [S_t, dl] concatenation
    calls = []
    def fake_randn_like(t):
        e = noise_for(t.shape[0], t.shape[1], 9000 + len(calls) + 17 * seed)
        calls.append(tuple(t.shape))
        return e.to(t)
    with mock.patch('torch.randn_like', side_effect = fake_randn_like):
        loss, breakdown, hiddens = model(batch, times = times, return_breakdown = True, return_hiddens = True)
    loss.backward()
    # structural ground truth straight from the reference's pack/route
    from transfusion_pytorch.modality_processing import get_processing_strategy
    with torch.no_grad(), mock.patch('torch.randn_like', side_effect = lambda t: torch.zeros_like(t)):
        samples = [[torch.tensor([model.sos_id]), *s, torch.tensor([model.eos_id])] for s in batch]
        samples = [[(0, p) if (torch.is_tensor(p) and p.is_floating_point()) else p for p in s] for s in samples]
        proc = get_processing_strategy('flat')(samples, times, model, need_axial_pos_emb = False, return_loss = True, return_embed = False)
    fx = dict(
        name = name, ctor = ctor, seed = seed, times = times, noise_shapes = calls,
        loss = loss.detach().double(), text_loss = breakdown.text.detach().double(), flow_losses = [f.detach().double() for f in breakdown.flow],
        modality_positions = proc.modality_positions, total_tokens = proc.total_tokens, text = proc.text.clone(),
        grads = grad_fingerprint(model),
    )
    embed = hiddens[-1].detach()
    if subsample_rows is not None:
        fx['embed_rows'] = subsample_rows
        fx['embed'] = embed[:, subsample_rows].clone()
    else:
        fx['embed'] = embed.clone()
        if keep_hiddens:
            fx['hiddens'] = [h.detach().clone() for h in hiddens[:-1]]
    torch.save(fx, os.path.join(GOLDEN, f'{name}.pt'))
    print(f'{name}: loss {loss.item():.6f} text {breakdown.text.item():.6f} flow {[round(f.item(), 6) for f in breakdown.flow]} '
          f'positions[0] {proc.modality_positions[0]} total_tokens {proc.total_tokens}')


def run_text_only(ref, name, ctor, text, seed, prompt_len = 16, gen_len = 40):
    torch.manual_seed(0)
    model = ref.Transfusion(**ctor)
    synth.fill_parameters_(model, seed = seed)
    model.eval()
    loss = model(text)
    loss.backward()
    gen = model.generate_text_only(text[:, :prompt_len], gen_len, temperature = 0.)
    with torch.no_grad():
        logits = model.forward_text(text[:, :-1], return_loss = False)
    fx = dict(name = name, ctor = ctor, seed = seed, loss = loss.detach().double(), grads = grad_fingerprint(model), generated = gen.clone(),
              prompt_len = prompt_len, gen_len = gen_len, logits_last = logits[:, -1].detach().clone())
    torch.save(fx, os.path.join(GOLDEN, f'{name}.pt'))
    print(f'{name}: loss {loss.item():.6f} generated[0,:8] {gen[0, :8].tolist()}')


def run_sampling(ref, name, ctor, seed):
    """sample_many on mixed prompts, greedy text, fixed init noise, forced modality at start (shrunk)."""
    torch.manual_seed(0)
    model = ref.Transfusion(**ctor)
    synth.fill_parameters_(model, seed = seed)
    model.eval()
    g = torch.Generator().manual_seed(77)
    dl = ctor['dim_latent']
    prompts = [
        torch.randint(0, ctor['num_text_tokens'], (9,), generator = g),
        (0, torch.randn(5, dl, generator = g)),
        None,
        [torch.randint(0, ctor['num_text_tokens'], (4,), generator = g), (0, torch.randn(7, dl, generator = g))],
    ]
    noise = torch.randn(16, dl, generator = g)
    kw = dict(max_length = 14, text_temperature = 0., cfg_scale = 3., modality_steps = 4, init_modality_noise = noise, force_modality_at_start = (0, (6,)),
              return_unprocessed_modalities = True)
    import copy
    out = model.sample_many(copy.deepcopy(prompts), **kw)
    fx = dict(name = name, ctor = ctor, seed = seed, prompts = prompts, noise = noise, kw = {k: v for k, v in kw.items() if k != 'init_modality_noise'}, samples = out)
    torch.save(fx, os.path.join(GOLDEN, f'{name}.pt'))
    for s in out:
        print('  sample:', [tuple(p.shape) if torch.is_tensor(p) else ('mod', p[0], tuple(p[1].shape)) for p in s])


def main():
    os.makedirs(GOLDEN, exist_ok = True)
    ref = load_reference()

    ctor = dict(num_text_tokens = 64, dim_latent = 32, modality_default_shape = (4,), transformer = dict(dim = 128, depth = 2, heads = 2))
    run_sampling(ref, 'sampling_small', ctor, seed = 8)
    if os.environ.get('GOLDEN_ONLY_SAMPLING'):
        return

    # (1) small single-modality, ragged, all hiddens kept
    ctor = dict(num_text_tokens = 64, dim_latent = 32, modality_default_shape = (4,), transformer = dict(dim = 128, depth = 2, heads = 2))
    batch = synth.small_batch(3, seed = 1, dim_latent = 32, text_vocab = 64)
    times = torch.rand(3, count_modalities(batch), generator = torch.Generator().manual_seed(5))
    run_interleaved(ref, 'small_one_modality', ctor, batch, times, seed = 1)

    # (2) small, depth 4 (two U-Net skips), two modality types, many short spans
    ctor = dict(num_text_tokens = 64, dim_latent = (32, 16), modality_default_shape = ((4,), (2,)), transformer = dict(dim = 128, depth = 4, heads = 4))
    batch = synth.config4_batch(2, seed = 2, total_len = 300, dims = (32, 16), text_vocab = 64)
    times = torch.rand(2, count_modalities(batch), generator = torch.Generator().manual_seed(6))
    run_interleaved(ref, 'small_two_modalities', ctor, batch, times, seed = 2)

    # (3) config 1: text-only pretrain shape (train_text_only.py), d=128 depth=2 heads=8
    ctor = dict(num_text_tokens = 256, transformer = dict(dim = 128, depth = 2))
    run_text_only(ref, 'config1_text_only', ctor, synth.text_batch(4, 257, seed = 3), seed = 3)

    # (4) config 2: the graded shape, b = 2 (CPU-feasible), embed subsampled
    ctor = dict(num_text_tokens = 256, dim_latent = 384, modality_default_shape = (256,), transformer = dict(dim = 512, depth = 8))
    batch = synth.config2_batch(2, seed = 4)
    times = synth.config2_times(2, seed = 4)
    rows = torch.tensor([0, 1, 5, 100, 205, 206, 207, 333, 461, 462, 500, 667, 668, 800, 923, 924, 1000, 1023])
    run_interleaved(ref, 'config2_b2', ctor, batch, times, seed = 4, subsample_rows = rows)


if __name__ == '__main__':
    main()
