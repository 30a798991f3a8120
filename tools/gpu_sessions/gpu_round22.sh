#!/bin/bash
# round 2, call 22: CML W=5 specialised kernel variants (64 registers / two buffers) vs generic
tag=${1:-r2O}
mkdir -p gpurun_out
for v in 1 2; do
( CF_STEP_FAST_W5=$v timeout 600 python bench.py --workload c2 --steps 30 --warmup 5 --no-cpu-baseline --no-other-configs --topk-users 0 > gpurun_out/${tag}_c2_v${v}.json 2> gpurun_out/${tag}_c2_v${v}.err; echo "bench c2 v5=$v rc=$?" )
python - <<PY
import json
j=json.loads(open('gpurun_out/${tag}_c2_v${v}.json').read().strip().splitlines()[-1])
r=j['roofline']
print('c2 v5=$v value %.3f G  ms %.3f  step %.3f apply %.3f count %.3f  whole_step_frac %.3f e2e %.3f G' % (j['value']/1e9, j['ms_per_step'], r['step_kernel_ms'], r['apply_kernel_ms'], r['count_kernel_ms'], r['whole_step_frac'], j['e2e']['value']/1e9))
PY
done
( CF_STEP_FAST_W5=2 timeout 900 python -m pytest tests/test_gpu_steps.py -x -q -k "specialised or cml" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" ); tail -3 gpurun_out/${tag}_pytest.log
