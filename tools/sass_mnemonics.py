#!/usr/bin/env python
"""Per-kernel counts of the Blackwell-specific / hot SASS instructions of libtfx_b200.so (evidence for tcgen05 / TMEM / TMA use):
    python tools/sass_mnemonics.py > profiles/r02_sass_mnemonics.txt"""
import collections, re, subprocess, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'transfusion_pytorch_b200', 'libtfx_b200.so')
out = subprocess.run(['cuobjdump', '-sass', lib], capture_output = True, text = True).stdout
WANT = ['UTCHMMA.2CTA', 'UTMALDG.2D.2CTA', 'UTCBAR.2CTA', 'UTCHMMA', 'UTMALDG', 'UTMAREDG', 'UTMASTG', 'LDTM', 'STTM', 'UTCBAR', 'SYNCS', 'FFMA2', 'FMUL2', 'FADD2', 'MUFU.EX2', 'HMMA', 'LDGSTS', 'REDG', 'UBLKCP']
print('# cuobjdump -sass transfusion_pytorch_b200/libtfx_b200.so : occurrences of the Blackwell-specific / hot instructions per kernel')
print('# UTCHMMA = tcgen05.mma (SS and TS form), UTMALDG = TMA load, UTMAREDG = TMA reduce-add, LDTM / STTM = tcgen05.ld / tcgen05.st, UTCBAR = tcgen05.commit,')
print('# .2CTA = cta_group::2 forms of the CTA-pair GEMMs (counted inside the plain mnemonic as well),')
print('# SYNCS = mbarrier, FFMA2/FMUL2/FADD2 = packed fp32x2, HMMA = legacy mma.sync (general attention path), LDGSTS = cp.async')
cur, cnt = None, collections.Counter()
def flush():
    if cur and cnt:
        print(cur[:150]); print('    ' + ', '.join(f'{k}={cnt[k]}' for k in WANT if cnt[k]))
for line in out.splitlines():
    m = re.match(r'\s*Function : (.*)', line)
    if m:
        flush(); cnt = collections.Counter()
        cur = subprocess.run(['c++filt', m.group(1)], capture_output = True, text = True).stdout.strip()
        continue
    for k in WANT:
        if re.search(r'\b' + re.escape(k) + r'\b', line) or (k in line and k in ('UTCHMMA', 'UTMALDG', 'UTMAREDG', 'LDTM', 'STTM', 'UTCBAR')):
            cnt[k] += 1
            break
flush()
