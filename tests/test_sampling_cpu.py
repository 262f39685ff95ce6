"""CPU: `sample_many` / `sample_one` host logic (state machine, midpoint ODE, CFG, prompt handling, the cached-KV
semantics of the reference) reproduces the reference's own `sample_many` output (tests/golden/sampling_small.pt)
when the engine is the fp32 oracle.  The same check runs on the GPU against the CUDA engine in test_sampling_gpu.py."""
import copy

import torch

from helpers import load_golden
from transfusion_pytorch_b200 import Transfusion, synth
from oracle.torch_reference import OracleEngine


def run(model, fx):
    return model.sample_many(copy.deepcopy(fx['prompts']), init_modality_noise = fx['noise'], **fx['kw'])


def compare(out, ref, text_exact = True, atol = 1e-4):
    assert len(out) == len(ref)
    for s, r in zip(out, ref):
        assert len(s) == len(r), ([type(p) for p in s], [type(p) for p in r])
        for a, b in zip(s, r):
            if torch.is_tensor(b):
                if text_exact:
                    assert torch.equal(a.cpu(), b), (a, b)
            else:
                assert a[0] == b[0] and a[1].shape == b[1].shape
                assert torch.allclose(a[1].float().cpu(), b[1], atol = atol, rtol = 1e-3), (a[1].float().cpu() - b[1]).abs().max()


def test_sample_many_matches_reference_with_oracle_engine():
    fx = load_golden('sampling_small')
    torch.manual_seed(0)
    model = Transfusion(**fx['ctor'])
    synth.fill_parameters_(model, seed = fx['seed'])
    model.eval()
    model._engine = OracleEngine(model)
    out = run(model, fx)
    compare(out, fx['samples'])
    one = model.sample_one(copy.deepcopy(fx['prompts'][3]), init_modality_noise = fx['noise'], **fx['kw'])
    compare([one], [fx['samples'][3]])


def test_sample_many_free_running_matches_reference_with_oracle_engine():
    """no forced modality: [som] tokens are SAMPLED (they never get a cache row and the modality takes their RoPE position, T.py:2337-2349,
    2408-2411), several text / modality rounds per sample, samples leave and re-enter the shared text loop at different times"""
    fx = load_golden('sampling_free')
    torch.manual_seed(0)
    model = Transfusion(**fx['ctor'])
    synth.fill_parameters_(model, seed = fx['seed'])
    model.eval()
    model._engine = OracleEngine(model)
    out = run(model, fx)
    n_prompt_mod = lambda p: 0 if (p is None or (torch.is_tensor(p) and not p.is_floating_point())) else (1 if not isinstance(p, list) else sum(not torch.is_tensor(q) or q.is_floating_point() for q in p))
    assert any(sum(not torch.is_tensor(p) for p in s) > n_prompt_mod(fx['prompts'][i]) for i, s in enumerate(fx['samples'])), 'fixture must contain a sampled modality'
    compare(out, fx['samples'])


def test_generate_text_only_and_forward_text_cache_with_oracle_engine():
    """host logic of the cached text path (T.py:2585-2707): greedy continuation equals the reference's bit for bit; `forward_text(cache=...)`
    token by token reproduces the un-cached logits (the reference's own cached-vs-uncached consistency test, tests/test_transfusion.py:559-662)"""
    fx = load_golden('config1_text_only')
    torch.manual_seed(0)
    model = Transfusion(**fx['ctor'])
    synth.fill_parameters_(model, seed = fx['seed'])
    model.eval()
    model._engine = OracleEngine(model)
    text = synth.text_batch(4, 257, seed = 3)
    gen = model.generate_text_only(text[:, :fx['prompt_len']], fx['gen_len'], temperature = 0.)
    assert torch.equal(gen, fx['generated'])
    with torch.no_grad():
        full = model.forward_text(text[:2, :12], return_loss = False)
        lg, cache = model.forward_text(text[:2, :8], return_loss = False, return_kv_cache = True)
        outs = [lg]
        for j in range(8, 12):
            lg, cache = model.forward_text(text[:2, j:j + 1], return_loss = False, cache = cache, return_kv_cache = True)
            outs.append(lg)
    assert cache[1] == 12
    assert torch.allclose(torch.cat(outs, dim = 1), full, atol = 1e-4, rtol = 1e-4)


def test_sample_many_config5_mid_matches_reference_with_oracle_engine():
    """config 5 of BASELINE.json at the size the GPU test uses (d 512, depth 8, 8 mixed prompts, forced 64 x 384 modality, 8 midpoint steps, cfg 3)"""
    fx = load_golden('config5_mid')
    torch.manual_seed(0)
    model = Transfusion(**fx['ctor'])
    synth.fill_parameters_(model, seed = fx['seed'])
    model.eval()
    model._engine = OracleEngine(model)
    out = run(model, fx)
    compare(out, fx['samples'])
