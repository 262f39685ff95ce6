"""CPU, world_size 2, gloo: the data-parallel step averages gradients across ranks and keeps replicas identical
(host-side logic of transfusion_pytorch_b200/data_parallel.py; the CUDA engine is replaced by the oracle engine)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR = '127.0.0.1', MASTER_PORT = str(port))
    dist.init_process_group('gloo', rank = rank, world_size = world)
    torch.set_num_threads(1)
    from transfusion_pytorch_b200 import Transfusion, synth
    from transfusion_pytorch_b200.data_parallel import DataParallelTrainer
    from oracle.torch_reference import OracleEngine
    torch.manual_seed(0)
    model = Transfusion(num_text_tokens = 64, dim_latent = 32, modality_default_shape = (4,), transformer = dict(dim = 128, depth = 2, heads = 2), prob_uncond = 0.)
    synth.fill_parameters_(model, seed = 7)
    model._engine = OracleEngine(model)
    tr = DataParallelTrainer(model, lr = 1e-3)
    batch = synth.small_batch(2, seed = 100 + rank, dim_latent = 32, text_vocab = 64)      # different shard per rank
    nm = max(sum(torch.is_tensor(p) and p.is_floating_point() for p in s) for s in batch)
    times = torch.rand(2, nm, generator = torch.Generator().manual_seed(rank))
    loss = tr.step(batch, times = times)
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        out.put((float(loss), bool(torch.equal(gathered[0], gathered[1])), float((flat - flat.mean()).abs().mean())))
    dist.destroy_process_group()


def test_two_rank_gloo_step_keeps_replicas_identical():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target = _worker, args = (r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    for p in procs: p.join(300)
    assert all(p.exitcode == 0 for p in procs)
    loss, same, _ = q.get(timeout = 10)
    assert same and loss == loss
