#!/bin/bash
# one GPU's share of configs[4] (12.5 M users x 10 M items, 625 M interactions, BPRMF W=1): specialised vs generic step kernel
tag=${1:-r2W}
mkdir -p gpurun_out
for gen in 0 1; do
( CF_STEP_GENERIC=$gen timeout 900 python bench.py --workload c5 --steps 20 --warmup 5 --no-cpu-baseline --no-other-configs --topk-users 0 > gpurun_out/${tag}_c5_gen${gen}.json 2> gpurun_out/${tag}_c5_gen${gen}.err; echo "bench c5 generic=$gen rc=$?" )
python - <<PY
import json
j=json.loads(open('gpurun_out/${tag}_c5_gen${gen}.json').read().strip().splitlines()[-1])
r=j['roofline']
print('c5 generic=$gen value %.3f G  ms %.3f  step %.3f apply %.3f count %.3f  frac %.3f whole_step_frac %.3f e2e %.3f G' % (j['value']/1e9, j['ms_per_step'], r['step_kernel_ms'], r['apply_kernel_ms'], r['count_kernel_ms'], r['frac'], r['whole_step_frac'], j['e2e']['value']/1e9))
PY
done
