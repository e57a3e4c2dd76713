#!/usr/bin/env python
"""Per-step launch table from an ncu gpu__time_duration launch list: tools/launch_table.py gpurun_out/x_launches.csv"""
import csv, re, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
L = []
for r in rows[1:]:
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1, 's': 1e3}.get(r[ui], 1e-6)
    L.append((re.sub(r'\(.*', '', r[ki]).replace('void ', ''), v))
idx = [n for n, (k, v) in enumerate(L) if k.startswith('k_status_pack')]
a, b = idx[-2] + 1, idx[-1] + 1
agg = collections.OrderedDict()
for k, v in L[a:b]:
    agg.setdefault(k, [0, 0.])
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v for k, v in L[a:b])
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-40s %3d %8.3f ms  %5.1f%%" % (k[:40], n, v, 100 * v / tot))
print("total %.3f ms, %d launches" % (tot, b - a))
