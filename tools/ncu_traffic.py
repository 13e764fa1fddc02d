#!/usr/bin/env python
"""DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum)
of the kernels in an ``ncu --set full`` capture of ONE bench step, written
as the JSON bench.py reads for ``roofline.traffic``.

Usage: ncu_traffic.py capture.ncu-rep "<workload tag>" label1 label2 ...
(labels in launch order, the names bench.py gives the kernels of a step)"""

import csv
import io
import json
import subprocess
import sys

rep, tag, labels = sys.argv[1], sys.argv[2], sys.argv[3:]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, units = rows[0], rows[1]
col = {k: i for i, k in enumerate(h)}

scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def val(r, k):
    return float(r[col[k]].replace(',', ''))*scale[units[col[k]]]


res = {}
for lab, r in zip(labels, rows[2:]):
    res[lab] = {
        'kernel': r[col['Kernel Name']],
        'dram_bytes': val(r, 'dram__bytes_read.sum') +
                      val(r, 'dram__bytes_write.sum'),
        'dram_read': val(r, 'dram__bytes_read.sum'),
        'dram_write': val(r, 'dram__bytes_write.sum'),
    }

print(json.dumps({'workload': tag, 'source': rep.split('/')[-1],
                  'kernels': res}, indent=1))
