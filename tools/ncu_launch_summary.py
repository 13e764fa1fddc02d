#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list:
per kernel name, launch count, mean duration and share of the captured
time.  Usage: ncu_launch_summary.py launches.csv [> summary.txt]"""

import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]
ki, vi, gi, bi = (h.index(c) for c in ('Kernel Name', 'Metric Value',
                                        'Grid Size', 'Block Size'))

agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) > vi:
        key = (r[ki], r[gi], r[bi])
        agg.setdefault(key, []).append(float(r[vi].replace(',', '')))

tot = sum(sum(v) for v in agg.values())
print(f'{"kernel":28s} {"grid":>16s} {"block":>14s} {"n":>4s} '
      f'{"mean_us":>10s} {"share":>7s}')
for (k, g, b), v in agg.items():
    print(f'{k[:28]:28s} {g:>16s} {b:>14s} {len(v):4d} '
          f'{sum(v)/len(v)/1e3:10.1f} {sum(v)/tot:7.3f}')
print(f'total captured: {tot/1e6:.3f} ms over {sum(map(len, agg.values()))} '
      'launches (cold-cache, serialised: compare shares, not absolutes)')
