#!/bin/bash
# A/B of the sampler / apply overlap knobs on the default bench workload (CML configs[1]); prints value, ms_per_step per variant
B="python bench.py --steps 60 --warmup 5 --no-other-configs --no-cpu-baseline --topk-users 0"
run() { echo "== $1"; env $1 $B 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); r = d['roofline']
        print('value %.4f G  ms %.4f  e2e %.4f G  k_step %.3f apply %.3f count %.3f' % (d['value'] / 1e9, d['ms_per_step'], d['e2e']['value'] / 1e9, r['step_kernel_ms'], r['apply_kernel_ms'], r['count_kernel_ms']))
"; }
run "CF_X=0"
run "CF_SAMPLE_PRIORITY=1"
run "CF_SAMPLE_PRIORITY=1 CF_APPLY_BLOCKS_PER_SM=6"
run "CF_APPLY_BLOCKS_PER_SM=6"
run "CF_SAMPLE_PRIORITY=1 CF_APPLY_BLOCKS_PER_SM=4"
run "CF_SAMPLE_PRIORITY=1 CF_SAMPLE_OVERLAP=1"
run "CF_SAMPLE_OVERLAP=0"
run "CF_X=0"
