"""Aggregate an ncu --csv launch list (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum) per kernel:
launches, total / average duration, DRAM bytes, achieved GB/s, fraction of the measured HBM peak."""
import csv
import json
import os
import re
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 7700.0
src = 'nominal 7.7 TB/s'
try:
    mp = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    for k in ('hbm_gbs', 'hbm_gbps'):
        if isinstance(mp.get(k), (int, float)):
            peak, src = float(mp[k]), 'MEASURED_PEAKS.json %s' % k
            break
except Exception:
    pass
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
iK, iM, iU, iV, iID = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Unit'), hdr.index('Metric Value'), hdr.index('ID')
per = OrderedDict()
for r in rows[1:]:
    kid = r[iID]
    d = per.setdefault(kid, {'name': r[iK]})
    v = float(r[iV].replace(',', ''))
    u = r[iU]
    scale = {'ns': 1e-9, 'nsecond': 1e-9, 'us': 1e-6, 'usecond': 1e-6, 'ms': 1e-3, 'msecond': 1e-3, 's': 1.0, 'second': 1.0,
             'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1.0)
    d[r[iM]] = v * scale
agg = OrderedDict()
for d in per.values():
    name = re.sub(r'^(void )?(dasr::)?', '', d['name'])
    name = re.sub(r'\(.*$', '', name)
    a = agg.setdefault(name, {'n': 0, 't': 0.0, 'b': 0.0, 'best': 0.0})
    t = d.get('gpu__time_duration.sum', 0.0)
    b = d.get('dram__bytes_read.sum', 0.0) + d.get('dram__bytes_write.sum', 0.0)
    a['n'] += 1
    a['t'] += t
    a['b'] += b
    if t > 5e-6:
        a['best'] = max(a['best'], b / t / 1e9)
print('HBM peak used: %.0f GB/s (%s)\n' % (peak, src))
print('| kernel | launches | total ms | avg us | DRAM MB / launch | achieved GB/s (all launches) | best launch GB/s | frac of HBM peak (best) |')
print('|---|---|---|---|---|---|---|---|')
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]['t']):
    gbs = a['b'] / a['t'] / 1e9 if a['t'] else 0.0
    print('| `%s` | %d | %.3f | %.1f | %.2f | %.0f | %.0f | %.2f |' % (name[:70], a['n'], a['t'] * 1e3, a['t'] / a['n'] * 1e6, a['b'] / a['n'] / 1e6, gbs, a['best'], a['best'] / peak))
