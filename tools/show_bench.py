#!/usr/bin/env python
"""Pretty-print a bench.py JSON line (default gpurun_out/bench.log)."""
import json, sys
path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/bench.log'
for l in open(path):
    if not l.startswith('{'):
        continue
    d = json.loads(l)
    print('value %.0f %s  ms/step %.3f  e2e %.0f  launches %s  clocks %s' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d.get('gpu_launches'), d.get('clocks')))
    r = d.get('roofline') or {}
    steps = max(1, min(d['steps'], 5))
    for k, v in sorted((r.get('kernels') or {}).items(), key=lambda kv: -kv[1]['ms']):
        if v['share'] > 0.004:
            print('  %-20s %4d launches %8.3f ms/step %5.1f%%' % (k, v['launches'], v['ms'] / steps, 100 * v['share']))
    print('  dominant %s (%s pipe): achieved %.2f TF (frac %.4f of tensor peak), executed %.2f TF = %.3f of its pipe; fp32 pipe %.1f TF; step algorithmic %.1f TF' % (
        r.get('kernel'), r.get('pipe'), r.get('achieved') or 0, r.get('frac') or 0, r.get('executed_tflops') or 0,
        r.get('executed_frac_of_pipe') or 0, (r.get('fp32_pipe') or {}).get('peak_tflops_measured') or 0, r.get('step_algorithmic_tflops') or 0))
    if d.get('cpu_baseline'):
        print('  cpu_baseline %.0f towers/s on %d cores' % (d['cpu_baseline']['value'], d['cpu_baseline']['cores']))
    for name, c in sorted((d.get('configs') or {}).items()):
        if c:
            print('  %s: %s' % (name, ', '.join('%s %s' % (k, ('%.4g' % v) if isinstance(v, float) else v) for k, v in c.items() if k != 'workload')))
