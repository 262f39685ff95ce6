#!/usr/bin/env python
"""Hottest SASS instructions of one kernel in an .ncu-rep captured with --import-source on.
usage: ncu_hot.py rep.ncu-rep <launch-skip> [top]"""
import csv, subprocess, sys
rep, skip = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--launch-skip', skip, '--launch-count', '1'], capture_output = True, text = True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:150])
h = rows[1]
isrc, ins, iex = h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
stall_cols = [(i, n) for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
data = []
for idx, r in enumerate(rows[2:]):
    try: data.append((int(r[ins]), int(r[iex]), r[isrc].strip(), idx, r))
    except (ValueError, IndexError): pass
tot = sum(d[0] for d in data)
print('total samples', tot, 'instructions', len(data))
for s, ex, src, idx, r in sorted(data, key = lambda d: -d[0])[:top]:
    st = sorted(((int(r[i] or 0), n) for i, n in stall_cols), reverse = True)[:2]
    print(f'{s:6d} {100 * s / tot:5.1f}%  ex={ex:8d} #{idx:5d} {src[:70]:70s} {st[0][1]}={st[0][0]} {st[1][1]}={st[1][0]}')
