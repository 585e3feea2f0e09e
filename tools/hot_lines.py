#!/usr/bin/env python
"""Hot CUDA source lines from `ncu -i REP --page source --print-source cuda --csv` output (file path in argv[1])."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
fname = None
hdr = None
out = []
for r in rows:
    if not r:
        continue
    if r[0] == 'File Name':
        fname = r[1].split('/')[-1]
        continue
    if r[0] == 'Line No':
        hdr = r
        continue
    if hdr and r[0].isdigit():
        d = dict(zip(hdr, r))
        s = d.get('# Samples', '0')
        if s.isdigit() and int(s) > 0:
            st = {k: int(d[k]) for k in hdr if k.startswith('stall_') and 'Not Issued' not in k and d.get(k, '') not in ('', '0')}
            out.append((int(s), fname, int(r[0]), d.get('Source', '').strip(), sorted(st.items(), key=lambda kv: -kv[1])[:3]))
tot = sum(o[0] for o in out)
print('total samples', tot)
for s, f, ln, src, st in out:
    if s > thr * tot:
        print('%6d %5.1f%% %s:%d  %-90s %s' % (s, 100 * s / tot, f, ln, src[:90], st))
