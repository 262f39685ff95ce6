"""GPU: sampling through the CUDA engine against the reference's `sample_many` output (tests/golden/sampling_small.pt).
Latents of the first decoded modality must agree within bf16 tolerance; greedy text before the first near-tie must be
identical; `sample_many` must equal `sample_one` per prompt (the reference's own self-consistency oracle,
tests/test_transfusion.py:758-808)."""
import copy

import pytest
import torch

from helpers import load_golden
from transfusion_pytorch_b200 import Transfusion, synth

pytestmark = pytest.mark.gpu


def build(fx):
    torch.manual_seed(0)
    model = Transfusion(**fx['ctor']).cuda()
    synth.fill_parameters_(model, seed = fx['seed'])
    return model.eval()


def test_sample_many_latents_and_structure_match_reference():
    fx = load_golden('sampling_small')
    model = build(fx)
    out = model.sample_many(copy.deepcopy(fx['prompts']), init_modality_noise = fx['noise'], **fx['kw'])
    for s, r in zip(out, fx['samples']):
        assert [torch.is_tensor(p) for p in s] == [torch.is_tensor(p) for p in r]
        mods_s = [p for p in s if not torch.is_tensor(p)]
        mods_r = [p for p in r if not torch.is_tensor(p)]
        dec_s, dec_r = mods_s[-1][1].float().cpu(), mods_r[-1][1]              # the modality decoded by the 3-eval-per-step ODE with CFG
        assert dec_s.shape == dec_r.shape
        assert (dec_s - dec_r).abs().max().item() < 5e-2 * dec_r.abs().max().item()
        assert torch.equal(s[0].cpu(), r[0])                                    # prompt text + forced [meta][shape][som] tokens


def test_sample_many_equals_sample_one():
    fx = load_golden('sampling_small')
    model = build(fx)
    many = model.sample_many(copy.deepcopy(fx['prompts']), init_modality_noise = fx['noise'], **fx['kw'])
    for i, prompt in enumerate(fx['prompts']):
        one = model.sample_one(copy.deepcopy(prompt), init_modality_noise = fx['noise'], **fx['kw'])
        for a, b in zip(one, many[i]):
            if torch.is_tensor(a):
                assert torch.equal(a.cpu(), b.cpu())
            else:
                assert torch.allclose(a[1].float(), b[1].float(), atol = 1e-5, rtol = 1e-5)     # batch-composition invariance of the kernels


def test_generate_modality_only_runs_and_is_finite():
    torch.manual_seed(0)
    model = Transfusion(num_text_tokens = 0, dim_latent = 32, modality_default_shape = (8,), transformer = dict(dim = 128, depth = 2, heads = 2)).cuda()
    synth.fill_parameters_(model, seed = 3)
    y = model.generate_modality_only(batch_size = 3, modality_steps = 4)
    assert y.shape == (3, 8, 32) and torch.isfinite(y).all()
    loss = model(torch.randn(3, 8, 32))
    loss.backward()
    assert torch.isfinite(loss) and model.engine.gflat.abs().sum() > 0
