#!/usr/bin/env python
"""DRAM bytes per launch of every kernel of ONE benchmarked step, keyed by the labels of bench.py's roofline table.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file step.csv \\
        python bench.py --no-graph --steps 1 --warmup 2 --no-cpu-baseline --dump-launches order.json
    python tools/ncu_traffic.py step.csv order.json profiles/r02_traffic.json

The last step of the capture (the launches between the last two fused-Adam kernels) is aligned with bench.py's launch-order dump: every C-ABI entry point
launches exactly one `tfx::` kernel, except the ones listed in EXTRA."""
import collections, csv, json, re, sys

EXTRA = {'attn_residual_bwd': 2, 'attn_residual_bwd_h16': 2, 'attn_residual_bwd2': 2}          # (the assembly-only call of bwd2 launches one kernel: handled below)          # main kernel + the parameter-gradient finish kernel

def main(csv_path, order_path, out_path):
    rows = list(csv.reader(open(csv_path, errors = 'replace')))
    hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r and 'Metric Name' in r)
    h = rows[hi]; iid, kn, mn, mv = h.index('ID'), h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value')
    launches = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv: continue
        d = launches.setdefault(r[iid], dict(name = r[kn]))
        try: d[r[mn]] = float(r[mv].replace(',', ''))
        except ValueError: pass
    seq = [d for d in launches.values() if 'tfx::' in d['name']]
    adam = [i for i, d in enumerate(seq) if 'adam_k' in d['name']]
    assert len(adam) >= 2, 'need at least two optimizer steps in the capture'
    step = seq[adam[-2] + 1: adam[-1] + 1]
    order = json.load(open(order_path))
    labels = order['launches']
    # the weight repack of a step is launched by the forward that FOLLOWS the optimizer: rotate so that both lists start at the same point
    count = lambda l: 1 if l.endswith('[assemble]') else EXTRA.get(re.sub(r'\[.*', '', l), 1)
    want = sum(count(l) for l in labels)
    assert want == len(step), f'alignment failed: bench lists {want} tfx kernels per step, the capture has {len(step)}'
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    i = 0
    for l in labels:
        n = count(l)
        for d in step[i:i + n]:
            agg[l][1] += d.get('dram__bytes_read.sum', 0.0) + d.get('dram__bytes_write.sum', 0.0)
            agg[l][2] += d.get('gpu__time_duration.sum', 0.0)
        agg[l][0] += 1
        i += n
    out = json.load(open(out_path)) if out_path and __import__('os').path.isfile(out_path) else {}
    out[order['key']] = {l: int(v[1] / v[0]) for l, v in agg.items()}
    out[order['key'] + ':ncu_us_per_launch'] = {l: round(v[2] / v[0] / 1e3, 1) for l, v in agg.items()}
    json.dump(out, open(out_path, 'w'), indent = 1, sort_keys = True)
    for l, v in sorted(agg.items(), key = lambda kv: -kv[1][2])[:20]:
        print(f'{l:42s} x{v[0]:3d}  {v[1] / v[0] / 1e6:9.1f} MB / launch   {v[2] / v[0] / 1e3:8.1f} us (ncu, serialised)')

if __name__ == '__main__':
    main(*sys.argv[1:4])
