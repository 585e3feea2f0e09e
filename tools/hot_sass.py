#!/usr/bin/env python
"""Hot SASS instructions of one kernel from `ncu -i REP --page source --csv` output (file path in argv[1])."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.006
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]
ia, isamp = hdr.index('Source'), hdr.index('# Samples')
data = []
for i, r in enumerate(rows[hi + 1:]):
    if len(r) <= isamp or r[0] == 'Address' or not r[isamp].isdigit():
        continue
    data.append((int(r[isamp]), i, r[ia].strip(), r))
tot = sum(d[0] for d in data)
print('total samples', tot, 'instructions', len(data))
for s, i, src, r in data:
    if s > tot * thr:
        d = dict(zip(hdr, r))
        st = {k: int(d[k]) for k in hdr if k.startswith('stall_') and 'Not Issued' not in k and d[k] not in ('', '0')}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print('%5d %5.1f%%  #%4d %-70s %s' % (s, 100 * s / tot, i, src[:70], top))
