#!/usr/bin/env python
"""Turn the ncu artefacts of tools/gpu_profile_r02*.sh (gpurun_out/) into what is committed under profiles/:
  profiles/<tag>_launches_summary.txt   per-kernel totals of the launch list (time, DRAM bytes, tensor-pipe activity)
  profiles/<tag>_ncu_full_summary.txt   selected metrics of every full capture + top stall reasons + SASS opcode histogram
  profiles/<tag>_ncu.json               what bench.py reads: per kernel tag DRAM bytes / launch and tensor-pipe activity, and the
                                        DRAM bytes of ONE training step (sum over the step's launches)
Usage: tools/summarize_ncu.py <tag>      (tag = r02)"""
import collections
import csv
import glob
import gzip
import json
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else 'r02'
STEPS_IN_RUN = 4        # bench.py --steps 1 --warmup 3 (+ 1 profiled step that launches with events: also a full step) -> counted below


def tag_of(name):
    """bench.py's profile tag of a demangled / base kernel name."""
    if 'k_wgrad_c' in name or 'k_wgrad_pair' in name:
        return 'k_wgrad_pair' if 'k_wgrad_pair' in name else 'k_wgrad_c'
    for k in ('k_edge_step_c', 'k_edge_dgrad_c', 'k_lin', 'k_gather_dsr_c', 'k_reduce_parts', 'k_skinny_c', 'k_seg_fix_c', 'k_edge_enc0_c'):
        if k in name:
            return k
    return name.split('(')[0].split('::')[-1]


def launches():
    rows = list(csv.reader(open('gpurun_out/launches_%s.csv' % tag)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr = rows[hi]
    ik, im, iv, iu, iid = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit'), hdr.index('ID')
    per = collections.OrderedDict()        # launch id -> dict
    for r in rows[hi + 1:]:
        if len(r) <= iv:
            continue
        d = per.setdefault(r[iid], {'name': r[ik]})
        v = float(r[iv].replace(',', ''))
        u = r[iu]
        if r[im] == 'gpu__time_duration.sum':
            v *= {'ns': 1e-6, 'nsecond': 1e-6, 'us': 1e-3, 'usecond': 1e-3, 'ms': 1.0, 'msecond': 1.0, 's': 1e3, 'second': 1e3}[u]
            d['ms'] = v
        elif r[im].startswith('dram__bytes'):
            v *= {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
            d['dram'] = d.get('dram', 0.0) + v
        elif 'pipe_tensor' in r[im]:
            d['tensor'] = v
    ls = list(per.values())
    # one training step = the launches between two consecutive k_edges_count launches (the edge build opens every step)
    starts = [i for i, d in enumerate(ls) if 'k_edges_count' in d['name']]
    step = ls[starts[-2]:starts[-1]] if len(starts) >= 2 else ls
    agg = collections.OrderedDict()
    for d in step:
        a = agg.setdefault(tag_of(d['name']), [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += d.get('ms', 0.0); a[2] += d.get('dram', 0.0); a[3] += d.get('tensor', 0.0) * d.get('ms', 0.0)
    tot_ms = sum(a[1] for a in agg.values()); tot_dram = sum(a[2] for a in agg.values())
    with open('profiles/%s_launches_summary.txt' % tag, 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active\n')
        f.write('#     --clock-control none -c 1000 --csv python bench.py --steps 1 --warmup 3 --no-configs\n')
        f.write('# ONE training step of the C2 workload (the launches between two edge builds; %d launches incl. torch optimiser kernels).\n' % len(step))
        f.write('# Per-launch times under ncu are cold-cache and serialised: compare SHARES.  tensor%% = time-weighted sm__pipe_tensor_cycles_active.\n')
        f.write('%-26s %8s %10s %8s %12s %9s\n' % ('kernel', 'launches', 'total_ms', 'share', 'dram_MB', 'tensor%'))
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('%-26s %8d %10.4f %7.2f%% %12.1f %8.1f%%\n' % (k, a[0], a[1], 100 * a[1] / tot_ms, a[2] / 1e6, a[3] / max(a[1], 1e-9)))
        f.write('%-26s %8d %10.4f %7s  %12.1f\n' % ('TOTAL', len(step), tot_ms, '', tot_dram / 1e6))
    return tot_dram, {k: dict(launches=a[0], ms=a[1], dram_bytes=a[2], tensor_active_pct=a[3] / max(a[1], 1e-9)) for k, a in agg.items()}


WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
BENCH_TAG = {  # full-capture kernel -> bench.py profile tag
    'k_wgrad_pair<1,': 'k_wgrad_c:step', 'k_wgrad_pair<0, 160': 'k_wgrad_c:enc', 'k_wgrad_c<0, 112': 'k_wgrad_c:node', 'k_gather_dsr_c': 'k_gather_dsr_c',
    'k_edge_step_c': 'k_edge_step_c',
    'k_edge_dgrad_c': 'k_edge_dgrad_c', 'k_lin<150, 150, 6153>': 'k_lin:enc_fwd', 'k_lin<150, 150, 640>': 'k_lin:enc_bwd', 'k_lin<100, 200': 'k_lin:node'}


def full():
    out = ['# ncu --set full --clock-control none --import-source on [--kernel-name-base demangled] -k regex:<kernel> -s 3 -c 1 python bench.py --steps 1 --warmup 3 --no-configs',
           '# one launch per kernel of the C2 training step; the .ncu-rep files are not committed (binary); numbers per launch', '']
    kern = {}
    for raw in sorted(glob.glob('gpurun_out/raw_%s_*.csv' % tag)):
        rows = list(csv.reader(open(raw)))
        if len(rows) < 3:
            continue
        hdr, units, r = rows[0], rows[1], rows[2]
        name = r[hdr.index('Kernel Name')]
        out.append('kernel: ' + name + '        [' + raw.split('/')[-1] + ']')
        vals = {}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                out.append('  %-72s %s %s' % (w, r[i], units[i]))
                try:
                    vals[w] = float(r[i].replace(',', ''))
                except ValueError:
                    pass
        stalls = []
        for h, v in zip(hdr, r):
            if 'issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h:
                try:
                    stalls.append((float(v), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
                except ValueError:
                    pass
        out.append('  top stall reasons (warps per issue): ' + ', '.join('%s %.2f' % (n, v) for v, n in sorted(stalls, reverse=True)[:6]))
        sass = raw.replace('raw_', 'sass_') + '.gz'
        try:
            srows = list(csv.reader(gzip.open(sass, 'rt')))
            sh = srows[1]
            iS, iE = sh.index('Source'), sh.index('Instructions Executed')
            ops = collections.Counter(); tot = 0
            for sr in srows[2:]:
                if len(sr) <= iE:
                    continue
                m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', sr[iS])
                try:
                    e = int(sr[iE])
                except ValueError:
                    continue
                if m:
                    ops[m.group(2)] += e; tot += e
            out.append('  executed warp instructions %d; by opcode: ' % tot + ', '.join('%s %.1f%%' % (o, 100.0 * c / tot) for o, c in ops.most_common(12)))
            tma = {o: c for o, c in ops.items() if o.startswith('UTMA') or o.startswith('UBLKCP') or o.startswith('UTC') or o in ('LDTM', 'STTM')}
            out.append('  tcgen05 / TMA instructions executed: ' + ', '.join('%s %d' % kv for kv in sorted(tma.items())))
        except (OSError, ValueError, IndexError):
            pass
        out.append('')
        for key, bt in BENCH_TAG.items():
            if key in name:
                kern[bt] = dict(kernel=name, dram_bytes_per_launch=(vals.get('dram__bytes_read.sum', 0.0) + vals.get('dram__bytes_write.sum', 0.0)) * 1e6,
                                tensor_active_pct=vals.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
                                issue_active_pct=vals.get('smsp__issue_active.avg.pct_of_peak_sustained_active'),
                                time_us_under_ncu=vals.get('gpu__time_duration.sum'))
    open('profiles/%s_ncu_full_summary.txt' % tag, 'w').write('\n'.join(out) + '\n')
    return kern


step_dram, per_kernel = launches()
kern = full()
json.dump({'source': 'profiles/%s_ncu_full_summary.txt, profiles/%s_launches_summary.txt' % (tag, tag), 'step_dram_bytes': step_dram,
           'step_kernels': per_kernel, 'kernels': kern}, open('profiles/%s_ncu.json' % tag, 'w'), indent=1)
print(open('profiles/%s_launches_summary.txt' % tag).read())
print(json.dumps(kern, indent=1))
