#!/usr/bin/env python
"""Summarise an ncu report (--page raw --csv) into the handful of metrics we track."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.avg.per_cycle_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg', 'smsp__inst_executed.sum']
idx = {h: i for i, h in enumerate(hdr)}
for r in rows[2:]:
    print('-----')
    for w in want:
        if w in idx:
            print(f'{w:72s} {r[idx[w]][:70]} {units[idx[w]]}')
    st = [(h, float(r[i])) for h, i in idx.items()
          if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio') and r[i]]
    st.sort(key=lambda x: -x[1])
    for h, v in st[:6]:
        print('   stall', h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), round(v, 2))
