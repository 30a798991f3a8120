#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): per-launch time, DRAM bytes, throughput %, occupancy, top stalls."""
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('==', d.get('Kernel Name', '?')[:90])
    for w in want:
        if w in d and d[w] not in ('', 'n/a'):
            print('   %-62s %s %s' % (w, d[w], units[hdr.index(w)]))
    st = []
    for k, v in d.items():
        if 'smsp__average_warp' in k and 'issue_stalled' in k and 'not_issued' not in k and v not in ('', 'n/a'):
            try:
                st.append((float(v.replace(',', '')), k.split('issue_stalled_')[1].split('_per_')[0]))
            except ValueError:
                pass
    print('   stalls/issue: ' + ', '.join('%s=%.2f' % (n, v) for v, n in sorted(st, reverse=True)[:7]))
