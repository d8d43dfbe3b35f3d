"""Summarise an ncu --csv launch list (gpu__time_duration.sum [+ dram bytes]) per kernel name / grid."""
import csv
import json
import sys
from collections import OrderedDict


def main(path, out_json=None):
    rows = [r for r in csv.reader(open(path, errors='ignore')) if len(r) > 10]
    hdr = rows[0]
    ik, im, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    iid = hdr.index('ID')
    per = OrderedDict()
    for r in rows[1:]:
        d = per.setdefault(r[iid], {'name': r[ik].split('(')[0]})
        v = float(r[iv].replace(',', ''))
        u = r[iu]
        if r[im].startswith('gpu__time_duration'):
            v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(u, 1.0)   # -> us
        else:
            v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1.0)
        d[r[im]] = v
    agg = OrderedDict()
    for d in per.values():
        a = agg.setdefault(d['name'], {'launches': 0, 'us': 0.0, 'dram': 0.0})
        a['launches'] += 1
        a['us'] += d.get('gpu__time_duration.sum', 0.0)
        a['dram'] += d.get('dram__bytes_read.sum', 0.0) + d.get('dram__bytes_write.sum', 0.0)
    tot = sum(a['us'] for a in agg.values())
    print('%-60s %8s %12s %7s %14s' % ('kernel', 'launches', 'total us', 'share', 'dram MB/launch'))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
        print('%-60s %8d %12.1f %6.1f%% %14.1f' % (k[:60], a['launches'], a['us'], 100 * a['us'] / tot, a['dram'] / a['launches'] / 1e6))
    print('total us: %.1f' % tot)
    if out_json:
        tcs = [v for k, v in agg.items() if 'conv_tc_kernel' in k or 'conv_tc2_kernel' in k]
        tc = {'launches': sum(v['launches'] for v in tcs), 'us': sum(v['us'] for v in tcs), 'dram': sum(v['dram'] for v in tcs)} if tcs else None
        if tc:
            json.dump({'kernel': 'dasr::conv_tc2_kernel + conv_tc_kernel', 'launches': tc['launches'],
                       'dram_bytes_per_launch_avg': tc['dram'] / tc['launches'], 'time_share': tc['us'] / tot,
                       'us_per_launch_avg_cold': tc['us'] / tc['launches'],
                       'source': 'ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum (tools/profile_forward.py)'},
                      open(out_json, 'w'), indent=1)


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
