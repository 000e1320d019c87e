#!/usr/bin/env python
"""Per-kernel SASS instruction counts of libsfft_b200.so (cuobjdump -sass): tensor / TMA / async-copy / local-memory mnemonics.
usage: python scripts/sass_summary.py > profiles/sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'sfft_b200', 'libsfft_b200.so')
out = subprocess.run(['cuobjdump', '-sass', so], stdout=subprocess.PIPE, text=True).stdout
demangle = lambda n: subprocess.run(['c++filt', n], stdout=subprocess.PIPE, text=True).stdout.strip()
WANT = ['DMMA', 'UBLKCP', 'UTMALDG', 'LDGSTS', 'SYNCS', 'STL', 'LDL', 'DFMA', 'DADD', 'DMUL', 'LDS', 'STS', 'BAR', 'USETMAXREG']
cur, cnt, tot, arch = None, collections.OrderedDict(), collections.Counter(), set()
for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
        cur = m.group(1); cnt[cur] = collections.Counter(); continue
    m = re.match(r'\s*arch = (\S+)', line)
    if m: arch.add(m.group(1))
    m = re.match(r'\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)', line)
    if m and cur:
        op = m.group(1)
        cnt[cur]['_all'] += 1
        for w in WANT:
            if op == w or op.startswith(w + '.') or (w in ('STL', 'LDL', 'LDS', 'STS', 'BAR') and op.startswith(w)):
                cnt[cur][w] += 1
print('# SASS summary of sfft_b200/libsfft_b200.so (cuobjdump -sass), architectures: %s' % ', '.join(sorted(arch)))
print('# kernel | instructions | ' + ' | '.join(WANT))
for k, c in sorted(cnt.items(), key=lambda kv: -kv[1]['_all']):
    if c['_all'] < 50: continue
    print('%-90s %6d  %s' % (demangle(k)[:90], c['_all'], '  '.join('%s=%d' % (w, c[w]) for w in WANT if c[w])))
    for w in WANT: tot[w] += c[w]
print('# totals: ' + '  '.join('%s=%d' % (w, tot[w]) for w in WANT))
