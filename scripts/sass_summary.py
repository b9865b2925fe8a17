#!/usr/bin/env python
"""Per-kernel SASS evidence for profiles/: counts of the mnemonics that prove TMA (UTMALDG), mbarrier use (SYNCS),
shared-memory gathers (LDS), global reductions (RED / ATOMG), packed fp32 (FFMA2 / FMUL2 / FADD2) and the absence of
tensor-core MMA in every kernel of libdfm.so.   python scripts/sass_summary.py [path/to/lib.so] > profiles/r2_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'multimodal-registration_b200', 'libdfm.so')
sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
names = {}
demangled = subprocess.run(['cu++filt'], input='\n'.join(set(re.findall(r'Function : (\S+)', sass))), capture_output=True, text=True)
KEYS = ['UTMALDG', 'UTMASTG', 'SYNCS', 'TLD4', 'LDS', 'STS', 'LDG', 'STG', 'REDG', 'ATOMG', 'ATOMS', 'SHFL', 'FFMA2', 'FMUL2', 'FADD2', 'FFMA', 'MMA']
rows = []
cur, cnt, total = None, None, 0
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        if cur:
            rows.append((cur, total, cnt))
        cur, cnt, total = m.group(1), collections.Counter(), 0
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and cur:
        total += 1
        op = m.group(1)
        for k in KEYS:
            if op.split('.')[0] == k or (k == 'MMA' and 'MMA' in op):
                cnt[k] += 1
if cur:
    rows.append((cur, total, cnt))
dem = dict(zip(sorted(set(r[0] for r in rows)), subprocess.run(['cu++filt'] + sorted(set(r[0] for r in rows)), capture_output=True, text=True).stdout.splitlines()))
agg = collections.OrderedDict()
for name, total, cnt in rows:
    base = re.sub(r'<.*', '', dem.get(name, name)).replace('void ', '').replace('dfm::', '')
    a = agg.setdefault(base, [0, 0, collections.Counter()])
    a[0] += 1
    a[1] += total
    a[2].update(cnt)
print('# %s: %d kernel instantiations in %d kernel templates (static SASS instruction counts, summed over instantiations)' % (os.path.basename(lib), len(rows), len(agg)))
print('%-28s %5s %8s ' % ('kernel', 'inst', 'SASS') + ' '.join('%7s' % k for k in KEYS))
for base, (n, total, cnt) in sorted(agg.items()):
    print('%-28s %5d %8d ' % (base[:28], n, total) + ' '.join('%7d' % cnt[k] for k in KEYS))
tot = collections.Counter()
for _, _, c in rows:
    tot.update(c)
print('%-28s %5d %8d ' % ('TOTAL', len(rows), sum(r[1] for r in rows)) + ' '.join('%7d' % tot[k] for k in KEYS))
