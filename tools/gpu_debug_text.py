import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from helpers import load_golden, grad_fingerprint
from transfusion_pytorch_b200 import Transfusion, synth
fx = load_golden('config1_text_only')
torch.manual_seed(0)
model = Transfusion(**fx['ctor']).cuda(); synth.fill_parameters_(model, seed = fx['seed']); model.eval()
text = synth.text_batch(4, 257, seed = 3)
loss = model(text); print('loss', loss.item(), fx['loss'].item(), 'grad_fn', loss.grad_fn)
loss.backward(); torch.cuda.synchronize()
print('gflat norm', model.engine.gflat.norm().item())
fp = grad_fingerprint((n, p.grad) for n, p in model.named_parameters() if p.grad is not None)
for k, v in fx['grads'].items():
    print(f'{k:60s} ours |g| {fp[k]["stats"][3].item():.4e} ref {v["stats"][3].item():.4e}  proj ours {fp[k]["stats"][2].item():.4e} ref {v["stats"][2].item():.4e}')
