#!/usr/bin/env python
"""Print the headline metrics and stall breakdown of every kernel in an
.ncu-rep file.  Usage: ncu_kernel_report.py report.ncu-rep"""

import csv
import io
import subprocess
import sys

out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[0]

keys = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_sector_hit_rate.pct',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__icc_request_hit_rate.pct', 'launch__registers_per_thread',
    'launch__grid_size', 'launch__block_size',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
    'smsp__inst_executed.sum',
    'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
    'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
    'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
    'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
]

for v in rows[2:]:
    d = dict(zip(h, v))
    unit = dict(zip(h, rows[1]))
    print(f"== {d.get('Kernel Name')}  grid {d.get('Grid Size')} block "
          f"{d.get('Block Size')}")
    for k in keys:
        if k in d:
            print(f'  {k:62s} {d[k]:>14s} {unit.get(k, "")}')

    ps = {k[33:]: float(x.replace(',', '')) for k, x in d.items()
          if k.startswith('smsp__pcsamp_warps_issue_stalled') and
          not k.endswith('not_issued') and x}
    tot = sum(ps.values()) or 1
    top = sorted(ps.items(), key=lambda t: -t[1])[:8]
    print('  stalls: ' + ', '.join(f'{k} {x/tot:.2f}' for k, x in top))
