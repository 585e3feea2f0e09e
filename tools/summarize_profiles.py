#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/) into the small text summaries
that are committed under profiles/.   Usage: tools/summarize_profiles.py <src-tag> <dst-tag>"""
import collections
import csv
import subprocess
import sys

src, dst = sys.argv[1], sys.argv[2]


def launches():
    rows = list(csv.reader(open('gpurun_out/launches_%s.csv' % src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr = rows[hi]
    ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= iv:
            continue
        name = r[ik].split('(')[0]
        v = float(r[iv].replace(',', ''))
        u = r[iu]
        v *= {'ns': 1e-6, 'nsecond': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 's': 1e3, 'second': 1e3}[u]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open('profiles/%s_launches_summary.txt' % dst, 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv python bench.py --steps 1 --warmup 3\n')
        f.write('# first 600 launches of the run; per-launch times are cold-cache and serialised: compare SHARES.\n')
        f.write('%-28s %8s %12s %8s\n' % ('kernel', 'launches', 'total_ms', 'share'))
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('%-28s %8d %12.4f %7.2f%%\n' % (k, a[0], a[1], 100 * a[1] / tot))


WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active']


def full():
    rep = 'gpurun_out/prof_%s.ncu-rep' % src
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = ['# ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 5 -c 2 python bench.py --steps 1 --warmup 3',
           '# source: %s (not committed: binary); numbers per launch' % rep, '']
    for r in rows[2:]:
        out.append('kernel: ' + r[hdr.index('Kernel Name')])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                out.append('  %-72s %s %s' % (w, r[i], units[i]))
        out.append('')
    srcp = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(srcp.splitlines()))
    hdr = rows[1]
    names = hdr[29:46]
    tot = [0] * len(names)
    ops = collections.Counter()
    nsamp = 0
    for r in rows[2:]:
        if r and r[0] == 'Kernel Name':
            break
        if len(r) < 63:
            continue
        for i, x in enumerate(r[29:46]):
            tot[i] += int(x)
        s = r[1].strip().split()
        op = (s[1] if s[0].startswith('@') else s[0]).split('.')[0]
        ops[op] += int(r[hdr.index('# Samples')])
        nsamp += int(r[hdr.index('# Samples')])
    S = sum(tot) or 1
    out.append('warp stall sampling, first launch (share of samples):')
    for n, t in sorted(zip(names, tot), key=lambda x: -x[1])[:10]:
        out.append('  %-26s %6.2f%%' % (n, 100.0 * t / S))
    out.append('samples by SASS opcode:')
    for op, s in ops.most_common(10):
        out.append('  %-10s %6.2f%%' % (op, 100.0 * s / max(nsamp, 1)))
    open('profiles/%s_ncu_full_summary.txt' % dst, 'w').write('\n'.join(out) + '\n')


if __name__ == '__main__':
    launches()
    full()
