#!/usr/bin/env python
"""bench.py - train tokens/sec (text+latent) at d=512 L=8 seq=1024 (BASELINE.json `metric`, configs[1]).

    python bench.py --gpus N --steps K --warmup W            # B200 arm (one process per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: the reference's CPU algorithm on the host cores

A "step" = forward + backward + gradient all-reduce (N > 1) + fused Adam (+ bf16 weight repack) over one synthetic batch of
`--batch` sequences x 1024 packed tokens per GPU (weak scaling).  Prints ONE JSON line on rank 0.

  value  : whole-job tokens/s with the packed batch already resident in HBM (device-timed, CUDA events, max over ranks); the step is a
           CUDA-graph replay (`DataParallelTrainer.step_packed`), `--no-graph` launches the same kernels eagerly
  e2e    : same metric through the public API - `DataParallelTrainer.step(list_of_samples)`: Python pack/route, H2D of the token metadata
           and of the latents from pinned host memory every step (copy stream), the step graph, and a D2H read of every step's loss
           (fetched one step late through a side stream so the host packs the next batch meanwhile)
  roofline: dominant kernel family of the step (by measured device time of an eager profiling pass), algorithmic FLOPs / measured time
           against the sustained measured bf16 peak; `roofline.kernels` lists the five largest kernel instances with their ncu DRAM traffic
  cpu_baseline: the oracle port of the reference algorithm (oracle/torch_reference.py, fp32, per-token conditioning, dense masks - the
           reference's cost structure) timed on this box's host cores on a bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CTOR = dict(num_text_tokens = 256, dim_latent = 384, modality_default_shape = (256,), transformer = dict(dim = 512, depth = 8))
SEQ = 1024
ALGO_TRAIN_FLOP_PER_TOKEN = 195.4e6            # 65.1 MFLOP/token forward x 3
METRIC = 'train tokens/sec (text+latent) at d=512 L=8 seq=1024'


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(path):
        p = json.load(open(path))
        return dict(hbm = p['hbm_gbs'], tf_burst = p['bf16_tflops'], tf_sustained = p.get('bf16_tflops_sustained', p['bf16_tflops']), src = 'measured')
    return dict(hbm = 6650., tf_burst = 1590., tf_sustained = 1400., src = 'fallback')


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons DURING the timed region"""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        super().__init__(daemon = True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-i', str(self.index)], capture_output = True, text = True, timeout = 5).stdout
                self.rows.append([c.strip() for c in out.strip().split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace('.', '').isdigit())
        reasons = set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        return dict(sm_mhz = sm[len(sm) // 2] if sm else None, sm_max_mhz = max(mx) if mx else None, reasons = sorted(reasons), samples = len(self.rows))


# --------------------------------------------------------------------------------------------- reference arm / cpu baseline
def cpu_port_tokens_per_s(batch: int, steps: int, warmup: int):
    """Times the oracle port (the checker, here only as the reported CPU baseline) - fwd + bwd + Adam on the host cores."""
    import torch
    from transfusion_pytorch_b200 import Transfusion, synth
    from oracle.torch_reference import OracleEngine
    cores = os.cpu_count() or 1
    torch.set_num_threads(min(cores, 64))
    cores = min(cores, 64)
    torch.manual_seed(0)
    model = Transfusion(**CTOR, prob_uncond = 0.)
    synth.fill_parameters_(model, seed = 0)
    model._engine = OracleEngine(model)
    opt = torch.optim.Adam(model.parameters(), lr = 1e-4)
    times_ = []
    for s in range(warmup + steps):
        b = synth.config2_batch(batch, seed = 500 + s)
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none = True)
        loss = model(b, times = synth.config2_times(batch, seed = s))
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if s >= warmup:
            times_.append(dt)
    ms = 1e3 * sum(times_) / len(times_)
    return batch * SEQ / (ms / 1e3), ms, cores


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    b = 2
    tps, ms, cores = cpu_port_tokens_per_s(b, max(args.steps, 1), min(args.warmup, 1))
    line = dict(impl = 'reference', metric = METRIC, value = tps, unit = 'tokens/s', n_gpus = args.gpus, steps = args.steps, warmup = args.warmup, ms_per_step = ms,
                higher_is_better = True, scaling = 'weak', vs_baseline = None, dtype = 'f32', data = 'synthetic',
                config = dict(workload = 'configs[1]: single-modality text+latent d=512 depth=8 dim_latent=384 seq=1024', global_batch = b, seq_len = SEQ, parallelism = 'cpu'),
                cpu_baseline = dict(value = tps, unit = 'tokens/s', cores = cores, kind = 'port', sample = f'{b} sequences x {SEQ} tokens per step, fwd+bwd+Adam, fp32, {args.steps} steps'),
                e2e = dict(value = tps, unit = 'tokens/s', h2d_bytes_per_step = 0, d2h_bytes_per_step = 0))
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------- B200 arm
def family_model(name, args_, eng, rb):
    """(family, algorithmic flops, algorithmic bytes) of one C-ABI launch, from its arguments."""
    a = args_
    if name == 'gemm_store':
        M, N, K = a[6], a[7], a[8]
        return 'gemm(tcgen05)', 2.0 * M * N * K, 0
    if name == 'gemm_qkvg':
        M, H, D = a[4], a[5], a[6]
        return 'gemm(tcgen05)', 2.0 * M * (3 * H * 64 + H) * D, 0
    if name == 'gemm_resid':
        M, N, K = a[7], a[8], a[9]
        Kr = eng.inner if K == eng.Ip else K
        return 'gemm(tcgen05)', 2.0 * M * N * Kr, 0
    if name == 'gemm_geglu':
        M, K = a[5], a[7]
        return 'gemm(tcgen05)', 2.0 * M * 2 * eng.inner * K, 0
    pairs = float(((rb.kv_limit.astype('int64') - (rb.cu[:-1].repeat(rb.seq_lens))) + 1).sum())
    if name == 'attn_fwd_tc':
        return 'attention', 4.0 * pairs * 64 * eng.H, 0
    if name == 'attn_fwd':
        return 'attention', 0, 0          # general kernel: returns at once when the tcgen05 path is active (flops credited to attn_fwd_tc)
    if name == 'attn_bwd_tc':
        return 'attention', 10.0 * pairs * 64 * eng.H, 0
    if name == 'attn_bwd':
        return 'attention', 0, 0
    return 'hbm-bound rows/elementwise', 0, 0


def log(*a):
    print(f'[bench {time.strftime("%H:%M:%S")}]', *a, file = sys.stderr, flush = True)


def run_b200_arm(args):
    import torch
    import torch.distributed as dist
    from transfusion_pytorch_b200 import Transfusion, synth
    from transfusion_pytorch_b200.data_parallel import DataParallelTrainer, AsyncScalar
    from transfusion_pytorch_b200.modality_processing import pack_batch

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get('NCCL_DEBUG', '').upper() == 'VERSION':     # NCCL's banner goes to stdout: keep stdout to the ONE JSON line
            os.environ['NCCL_DEBUG'] = 'WARN'
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id = torch.device('cuda', local))
    dev = torch.device('cuda', local)
    B = args.batch
    torch.manual_seed(0)
    model = Transfusion(**CTOR).to(dev)                      # prob_uncond = 0.1 (reference default), train mode
    synth.fill_parameters_(model, seed = 0)
    model.train()
    trainer = DataParallelTrainer(model, lr = 1e-4, cuda_graph = not args.no_graph)
    eng = model.engine
    eng.ensure_attached()

    POOL = 4
    host_batches = [synth.config2_batch(B, seed = 1000 * rank + i) for i in range(POOL)]
    host_batches = [[[p.pin_memory() if p.is_floating_point() else p for p in s] for s in b] for b in host_batches]
    host_times = [synth.config2_times(B, seed = 1000 * rank + i) for i in range(POOL)]

    # ---- device-resident variant: packed descriptors + latents already in HBM
    packed = []
    for b, t in zip(host_batches, host_times):
        samples = [[torch.tensor([model.sos_id]), *s, torch.tensor([model.eos_id])] for s in b]
        rb = pack_batch(samples, t, model, return_loss = True, return_embed = False)
        lat = model._latents_to_device(rb)
        eng.upload(rb)
        packed.append((rb, lat))
    assert packed[0][0].M == B * SEQ
    log('packed', POOL, 'batches; M =', packed[0][0].M)

    profiling = [False]                                      # roofline pass: every rank launches eagerly (same collectives on all ranks)
    def step_resident(i):
        rb, lat = packed[i % POOL]
        if trainer.cuda_graph and not profiling[0]:
            return trainer.step_packed(rb, lat)              # CUDA-graph replay of the step (after two eager steps of this shape)
        eng.zero_grad()
        loss = model.forward_packed(rb, lat)
        if world > 1 and trainer.overlap:
            # same overlap machinery as DataParallelTrainer.step
            bounds = trainer._bucket_bounds(eng)
            if trainer.comm_stream is None:
                trainer.comm_stream = torch.cuda.Stream()
            hi = [trainer._tail]
            def cb(l):
                lo = bounds[l]
                ev = torch.cuda.Event(); ev.record()
                with torch.cuda.stream(trainer.comm_stream):
                    trainer.comm_stream.wait_event(ev)
                    dist.all_reduce(eng.gflat[lo:hi[0]])
                hi[0] = lo
            eng._bucket_cb = cb
            loss.backward()
            eng._bucket_cb = None
            ev = torch.cuda.Event(); ev.record()
            with torch.cuda.stream(trainer.comm_stream):
                trainer.comm_stream.wait_event(ev)
                if hi[0] > 0:
                    dist.all_reduce(eng.gflat[:hi[0]])
                if trainer._tail < eng.gflat.numel():
                    dist.all_reduce(eng.gflat[trainer._tail:])
            torch.cuda.current_stream().wait_stream(trainer.comm_stream)
        else:
            loss.backward()
        eng.adam_step(lr = 1e-4, grad_scale = 1.0 / world, zero_grads = True)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing = True), torch.cuda.Event(enable_timing = True)
        e0.record()
        t_host = time.perf_counter()
        for i in range(steps):
            fn(i)
        timed.host_ms = 1e3 * (time.perf_counter() - t_host) / steps      # CPU time to enqueue one step (no sync inside)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device = dev)
        if world > 1:
            dist.all_reduce(ms, op = dist.ReduceOp.MAX)
        return ms.item()

    for i in range(args.warmup):
        step_resident(i)
        torch.cuda.synchronize(); log('warmup step', i, 'done')
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler: sampler.start()
    l0 = eng.ops.launches
    ms_total = timed(step_resident, args.steps)
    host_enqueue_ms = timed.host_ms
    launches = eng.ops.launches - l0
    if sampler:
        sampler.stop_flag = True
    ms_step = ms_total / args.steps
    log('resident ms/step', ms_step)
    value = world * B * SEQ / (ms_step / 1e3)

    # ---- end to end through the public API (pack/route + H2D + D2H every step)
    # Every step: Python pack/route of host samples, H2D of that step's inputs from pinned memory, fwd + bwd + optimizer, and a D2H
    # read of a loss.  The loss that is read inside step i is the one of step i-1 (asynchronous logging: the value is fetched while step i
    # runs on the device, so the host packs step i+1 instead of idling); the last loss is read before the timed region closes.
    h2d = [0]
    pending = [None]
    host_e2e = []
    def step_e2e(i):
        b, t = host_batches[i % POOL], host_times[i % POOL]
        t_h = time.perf_counter()
        loss = trainer.step(b, times = t)
        host_e2e.append(1e3 * (time.perf_counter() - t_h))
        rb = model._last_batch
        h2d[0] = rb.dev.get('h2d_bytes', 0) + getattr(rb, 'latent_h2d_bytes', 0)
        prev, pending[0] = pending[0], AsyncScalar(loss)      # D2H copy of this step's loss, on a side stream
        return prev.value() if prev is not None else None    # ... read one step later: waits for step i-1 only
    def e2e_loop(i):
        step_e2e(i)
        if i == e2e_loop.last:
            pending[0].value(); pending[0] = None              # drain: the final step's loss is read inside the timed region too
    for i in range(min(args.warmup, 3)):
        step_e2e(i)
    pending[0].value(); pending[0] = None
    e2e_steps = max(3, min(args.steps, 10))
    e2e_loop.last = e2e_steps - 1
    ms_e2e = timed(e2e_loop, e2e_steps) / e2e_steps
    e2e_value = world * B * SEQ / (ms_e2e / 1e3)
    log('e2e ms/step', ms_e2e)

    # ---- per-kernel-family device time of one step (profiling pass, not part of the reported throughput)
    # every rank runs the step (it contains the gradient all-reduce); only rank 0 records per-launch events
    roof = None
    profiling[0] = True
    if rank == 0:
        eng.ops.timing = {}
    step_resident(0)
    barrier()
    if rank == 0:
        fam, inst = {}, {}
        rb = packed[0][0]
        for name, recs in eng.ops.timing.items():
            for e0, e1, a in recs:
                f, fl, by = family_model(name, a, eng, rb)
                ms = e0.elapsed_time(e1)
                d = fam.setdefault(f, dict(ms = 0., flops = 0., launches = 0))
                d['ms'] += ms; d['flops'] += fl; d['launches'] += 1
                if fl > 0:                                    # per kernel instance (entry point + problem shape)
                    label = name + (f'[M={a[6]},N={a[7]},K={a[8]}]' if name == 'gemm_store' else '')
                    k = inst.setdefault(label, dict(ms = 0., flops = 0., launches = 0))
                    k['ms'] += ms; k['flops'] += fl; k['launches'] += 1
        eng.ops.timing = None
        pk = peaks()
        tot = sum(d['ms'] for d in fam.values())
        top = max((f for f in fam if fam[f]['flops'] > 0), key = lambda f: fam[f]['ms'])
        ach = fam[top]['flops'] / (fam[top]['ms'] / 1e3) / 1e12
        roof = dict(bound = 'tensor', kernel = top, achieved = ach, peak = pk['tf_sustained'], unit = 'TFLOP/s', frac = ach / pk['tf_sustained'], traffic = None,
                    peak_source = pk['src'] + ' (sustained bf16 GEMM)', share_of_step = fam[top]['ms'] / tot,
                    families = {f: dict(ms = round(d['ms'], 3), share = round(d['ms'] / tot, 3), launches = d['launches'],
                                        tflops = round(d['flops'] / (d['ms'] / 1e3) / 1e12, 1) if d['flops'] else None) for f, d in fam.items()},
                    whole_step_tflops = value * ALGO_TRAIN_FLOP_PER_TOKEN / 1e12 / world, whole_step_frac = value * ALGO_TRAIN_FLOP_PER_TOKEN / 1e12 / world / pk['tf_sustained'])
        # the five kernel instances with the largest share of the step: algorithmic FLOPs per launch / average launch duration (CUDA events),
        # DRAM traffic per launch from the committed ncu captures (measured at batch 32; null for other batch sizes / kernels)
        tpath = os.path.join(ROOT, 'profiles', 'r01_traffic.json')
        tmap = json.load(open(tpath)) if os.path.isfile(tpath) else {}
        top = sorted(inst.items(), key = lambda kv: -kv[1]['ms'])[:5]
        roof['kernels'] = [dict(kernel = lbl, launches = k['launches'], us_per_launch = round(1e3 * k['ms'] / k['launches'], 1), share_of_step = round(k['ms'] / tot, 3),
                                achieved = round(k['flops'] / (k['ms'] / 1e3) / 1e12, 1), frac = round(k['flops'] / (k['ms'] / 1e3) / 1e12 / pk['tf_sustained'], 3),
                                traffic = (tmap.get(lbl) if B == 32 else None)) for lbl, k in top]

    if rank == 0:
        clocks = sampler.summary() if sampler else None
        cpu = None
        log('roofline pass done')
        if world == 1 and not args.no_cpu_baseline:
            tps, ms_cpu, cores = cpu_port_tokens_per_s(2, 2, 1)
            cpu = dict(value = tps, unit = 'tokens/s', cores = cores, kind = 'port', sample = f'2 sequences x {SEQ} tokens per step (fwd+bwd+Adam, fp32), 2 timed steps after 1 warm-up')
        line = dict(metric = METRIC, value = value, unit = 'tokens/s', n_gpus = world, steps = args.steps, warmup = args.warmup, ms_per_step = ms_step, higher_is_better = True,
                    scaling = 'weak', vs_baseline = None, dtype = 'bf16', data = 'synthetic',
                    config = dict(workload = 'configs[1]: single-modality text+latent d=512 depth=8 dim_latent=384 seq=1024', global_batch = world * B, per_gpu_batch = B,
                                  seq_len = SEQ, parallelism = f'dp{world}', optimizer = 'fused Adam', launch = 'cuda graph replay' if trainer.cuda_graph else 'eager', l2 = 'per-step working set (>10 GB of activations) is far larger than the 126 MB L2; 4 rotating input batches'),
                    e2e = dict(value = e2e_value, unit = 'tokens/s', ms_per_step = ms_e2e, h2d_bytes_per_step = int(h2d[0]), d2h_bytes_per_step = 4,
                               loss_read = 'every step, deferred by one step (asynchronous logging)',
                               host_ms_per_step = round(sum(host_e2e[-e2e_steps:]) / e2e_steps, 3)),
                    gpu_launches = int(launches), host_enqueue_ms_per_step = round(host_enqueue_ms, 3), clocks = clocks, roofline = roof, cpu_baseline = cpu)
        print(json.dumps(line))
    if world > 1:
        # captured graphs hold NCCL work: drop them, drain the device, leave together.  The process then exits without running the
        # process-group destructor (observed to hang after graph-captured collectives); every rank has already passed the barrier.
        trainer._graphs.clear()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type = int, default = 1)
    ap.add_argument('--steps', type = int, default = 10)
    ap.add_argument('--warmup', type = int, default = 3)
    ap.add_argument('--batch', type = int, default = 128, help = 'sequences (x1024 tokens) per GPU per step (swept 32 / 64 / 128 on B200: 2.00 / 2.12 / 2.18 M tokens/s)')
    ap.add_argument('--impl', default = 'b200', choices = ['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action = 'store_true')
    ap.add_argument('--no-graph', action = 'store_true', help = 'eager kernel launches instead of CUDA-graph replay (N = 1)')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == '__main__':
    main()
