"""localises the cached-vs-uncached forward_text mismatch: per (sample, position) max |diff| of the logits"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from helpers import load_golden
from transfusion_pytorch_b200 import Transfusion, synth
fx = load_golden('config1_text_only')
torch.manual_seed(0)
model = Transfusion(**fx['ctor']).cuda(); synth.fill_parameters_(model, seed = fx['seed']); model.eval()
text = synth.text_batch(3, 300, seed = 9).cuda()
with torch.no_grad():
    full = model.forward_text(text[:, :200], return_loss = False).float().clone()
    lg, cache = model.forward_text(text[:, :150], return_loss = False, return_kv_cache = True)
    outs = [lg.float().clone()]
    for j in range(150, 200):
        lg, cache = model.forward_text(text[:, j:j + 1], return_loss = False, cache = cache, return_kv_cache = True)
        outs.append(lg.float().clone())
got = torch.cat(outs, dim = 1)
d = (got - full).abs().amax(dim = -1)
torch.set_printoptions(linewidth = 250, precision = 3)
print('prefill part max', d[:, :150].max().item(), 'decode part max', d[:, 150:].max().item())
print('per-sample decode diffs'); print(d[:, 150:])
print('prefill diffs > 1e-2 at', (d[:, :150] > 1e-2).nonzero().tolist()[:40])
print('cap', cache[0].cache.cap, 'len', cache[0].length)
