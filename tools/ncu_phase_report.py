#!/usr/bin/env python
"""Per-phase profile of a kernel from an .ncu-rep captured with
``--set full --import-source on``: the SASS listing is cut at every
``BAR.SYNC`` and the stall samples, executed instructions, FP64
instructions and shared-memory wavefronts are summed per segment.

Usage: ncu_phase_report.py report.ncu-rep [kernel-name-regex]"""

import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else '.')

out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'],
                     capture_output=True, text=True).stdout

# The CSV holds one table per kernel, each introduced by a "Kernel Name" row
tables, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if row and row[0] == 'Kernel Name':
        cur = {'name': row[1], 'rows': []}
        tables.append(cur)
    elif cur is not None and row:
        cur['rows'].append(row)

for t in tables:
    if not pat.search(t['name']):
        continue

    hdr, rows = t['rows'][0], t['rows'][1:]
    col = {h: i for i, h in enumerate(hdr)}
    num = lambda r, k: float(r[col[k]] or 0) if k in col else 0.0

    segs, seg = [], dict(n=0, samples=0, inst=0, fp64=0, smem=0, lds=0,
                         stg=0, first=None)
    for r in rows:
        sass = r[col['Source']].strip()
        seg['n'] += 1
        seg['samples'] += num(r, '# Samples')
        ie = num(r, 'Instructions Executed')
        seg['inst'] += ie
        seg['smem'] += num(r, 'L1 Wavefronts Shared')
        if re.match(r'(@!?U?P\d+\s+)?D(FMA|MUL|ADD)', sass):
            seg['fp64'] += ie
        if re.match(r'(@!?U?P\d+\s+)?(LDS|STS)', sass):
            seg['lds'] += ie
        if re.match(r'(@!?U?P\d+\s+)?STG', sass):
            seg['stg'] += ie
        if 'BAR.SYNC' in sass or sass.startswith('EXIT'):
            segs.append(seg)
            seg = dict(n=0, samples=0, inst=0, fp64=0, smem=0, lds=0, stg=0,
                       first=None)
    if seg['n']:
        segs.append(seg)

    tot = sum(s['samples'] for s in segs) or 1
    print(f"== {t['name']}: {len(rows)} SASS instructions, "
          f"{int(tot)} samples")
    print(' seg  #sass  samples%  warp-inst(M)  fp64(M)  lds/sts(M)  '
          'stg(M)  smem-wavefronts(M)')
    for i, s in enumerate(segs):
        print(f"{i:4d} {s['n']:6d} {100*s['samples']/tot:8.1f} "
              f"{s['inst']/1e6:12.1f} {s['fp64']/1e6:8.1f} "
              f"{s['lds']/1e6:10.1f} {s['stg']/1e6:7.1f} {s['smem']/1e6:12.1f}")
