#!/bin/bash
# batch-size sweep helper (GPU box): prints B, update mode, Gtriples/s, kernel GB/s, frac
for U in ${3:-sync hogwild}; do
for B in ${4:-4096 16384 65536 262144 1048576}; do
  python bench.py --workload ${1:-c2} --batch $B --steps ${2:-20} --warmup 3 --no-cpu-baseline --topk-users 0 --update $U 2>&1 | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read())
r=j['roofline']
print('%s B=%d %s value=%.3fG/s e2e=%.3fG/s ms/step=%.3f kernel_ms=%.3f count_ms=%.3f apply_ms=%.3f GB/s=%.0f frac=%.3f' % ('$1', $B, '$U', j['value']/1e9, j['e2e']['value']/1e9, j['ms_per_step'], r['kernel_ms_per_launch'], r['count_kernel_ms_per_launch'], r['apply_kernel_ms_per_launch'], r['achieved'], r['frac']))"
done; done
