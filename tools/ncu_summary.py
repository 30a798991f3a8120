#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): per-launch time, DRAM bytes, throughput %, occupancy, top stalls."""
import csv
import subprocess
import sys

rep = sys.argv[1]
# --traffic KEY REGEX[,REGEX..] SOURCE: also record the summed DRAM bytes of the FIRST launch matching each kernel regex under
# KEY in profiles/traffic.json (what bench.py's roofline.traffic reads; KEY = "<workload>|B=<pairs>|<optimizer>|<update>")
traffic = None
if '--traffic' in sys.argv:
    k = sys.argv.index('--traffic')
    traffic = dict(key=sys.argv[k + 1], kernels=sys.argv[k + 2].split(','), source=sys.argv[k + 3])
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

if traffic:
    import json
    import os
    import re

    def gb(d, name):
        v, u = float(d[name].replace(',', '')), units[hdr.index(name)]
        return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Tbyte': 1e12}[u]
    kern, total = {}, 0.0
    for pat in traffic['kernels']:
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            if re.search(pat, d.get('Kernel Name', '')):
                kern[d['Kernel Name'][:60]] = dict(read=gb(d, 'dram__bytes_read.sum'), write=gb(d, 'dram__bytes_write.sum'),
                                                   ms=d.get('gpu__time_duration.sum'))
                total += gb(d, 'dram__bytes_read.sum') + gb(d, 'dram__bytes_write.sum')
                break
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'traffic.json')
    j = json.load(open(path)) if os.path.exists(path) else {}
    j[traffic['key']] = dict(bytes_per_launch=total, kernels=kern, source=traffic['source'])
    json.dump(j, open(path, 'w'), indent=1)
    print('traffic[%s] = %.3f GB' % (traffic['key'], total / 1e9))
