#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and (optionally) the sequence."""
import csv, collections, gzip, re, sys

def load(path):
    rows = list(csv.reader(gzip.open(path, 'rt', errors = 'replace') if path.endswith('.gz') else open(path, errors = 'replace')))
    hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    h = rows[hi]; kn, mv, gs = h.index('Kernel Name'), h.index('Metric Value'), h.index('Grid Size')
    mn = h.index('Metric Name') if 'Metric Name' in h else None
    out = []
    for r in rows[hi + 1:]:
        if len(r) <= mv: continue
        if mn is not None and r[mn] != 'gpu__time_duration.sum': continue      # (captures with several metrics per launch: keep the duration rows)
        try: v = float(r[mv].replace(',', ''))
        except ValueError: continue
        name = re.sub(r'^void ', '', r[kn]); name = re.sub(r'\(.*', '', name)
        out.append((name, r[gs], v / 1e3))
    return out

if __name__ == '__main__':
    seq = load(sys.argv[1])
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, g, us in seq:
        agg[n][0] += 1; agg[n][1] += us
    tot = sum(v[1] for v in agg.values())
    for k, v in sorted(agg.items(), key = lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 45]:
        print(f'{v[1]:10.1f} us {v[0]:5d} x {v[1] / v[0]:8.1f} us {100 * v[1] / tot:5.1f}%  {k[:100]}')
    print(f'total {tot / 1e3:.3f} ms over {len(seq)} launches')
    if len(sys.argv) > 3:
        for s in seq: print(s)
