"""Summarise an ncu pass over tools/hbm_bench.py into a markdown table (per kernel: launches, mean duration, DRAM bytes, DRAM
throughput):   python tools/hbm_ncu_summary.py gpurun_out/r2_hbm_ncu.csv"""
import csv, sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith('==')]
hdr = rows[0]
ik, im, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
iid = hdr.index('ID')
per = OrderedDict()
for r in rows[1:]:
    if len(r) != len(hdr):
        continue
    v = float(r[iv].replace(',', ''))
    u = r[iu]
    if r[im] == 'gpu__time_duration.sum':
        v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}.get(u, 1.0)
    elif r[im].startswith('dram__bytes'):
        v *= {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(u, 1.0)
    per.setdefault((r[iid], r[ik].split('(')[0].split('::')[-1]), {})[r[im]] = v
agg = OrderedDict()
for (_, k), m in per.items():
    if m.get('gpu__time_duration.sum', 0) < 20:        # memsets, table uploads
        continue
    a = agg.setdefault(k, [])
    a.append(m)
print('| kernel | launches | us | DRAM read MB | DRAM write MB | dram throughput % of peak |\n|---|---|---|---|---|---|')
for k, ms in agg.items():
    f = lambda name: sum(m.get(name, 0.0) for m in ms) / len(ms)
    print('| `%s` | %d | %.1f | %.1f | %.1f | %.1f |' % (k, len(ms), f('gpu__time_duration.sum'), f('dram__bytes_read.sum'),
                                                      f('dram__bytes_write.sum'), f('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')))
