#!/usr/bin/env python
"""Static counts of the tcgen05 / TMEM / TMA / mbarrier SASS mnemonics per kernel: cuobjdump -sass libspwgnn.so | tools/sass_mnemonics.py > profiles/<tag>_sass_mnemonics.txt"""
import collections
import re
import subprocess
import sys

fn = None
per = collections.OrderedDict()
for l in sys.stdin:
    m = re.search(r'Function : (\S+)', l)
    if m:
        fn = m.group(1)
        per[fn] = collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
    if m and fn:
        per[fn][m.group(2).split('.')[0]] += 1
keys = ['UTCHMMA', 'UTCBAR', 'UTMALDG', 'UBLKCP', 'LDTM', 'STTM', 'UTCATOMSWS', 'LDGSTS', 'SYNCS', 'SHFL', 'USETMAXREG']
print('# cuobjdump -sass spwgnn_b200/libspwgnn.so: static instruction counts of the tcgen05 / TMEM / TMA / mbarrier mnemonics per kernel')
print('# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UTCATOMSWS = tcgen05.alloc, UTMALDG = cp.async.bulk.tensor')
print('# (TMA tile load through a tensor map), UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier ops, USETMAXREG = setmaxnreg, LDGSTS = cp.async')
print()
print('%-64s %6s ' % ('kernel', 'instr') + ' '.join('%7s' % k[:7] for k in keys))
for f, c in per.items():
    if not any(c[k] for k in keys[:7]):
        continue
    name = subprocess.run(['c++filt', f], capture_output=True, text=True).stdout.strip().split('(')[0].replace('void ', '').replace('spw::', '')
    print('%-64s %6d ' % (name[-64:], sum(c.values())) + ' '.join('%7d' % c[k] for k in keys))
