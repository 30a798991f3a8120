#!/bin/bash
# round 2, call 21: specialised step kernels (cf_step_fast.cu) vs the generic one: parity tests + A/B timing
tag=${1:-r2N}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_steps.py tests/test_gpu_e2e.py -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" ); tail -5 gpurun_out/${tag}_pytest.log
for wl in c2-bpr c2; do
for gen in 0 1; do
( CF_STEP_GENERIC=$gen timeout 600 python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu-baseline --no-other-configs --topk-users 0 > gpurun_out/${tag}_${wl}_gen${gen}.json 2> gpurun_out/${tag}_${wl}_gen${gen}.err; echo "bench $wl generic=$gen rc=$?" )
python - <<PY
import json
j=json.loads(open('gpurun_out/${tag}_${wl}_gen${gen}.json').read().strip().splitlines()[-1])
r=j['roofline']
print('$wl generic=$gen value %.3f G  ms %.3f  step %.3f apply %.3f count %.3f  whole_step_frac %.3f e2e %.3f G' % (j['value']/1e9, j['ms_per_step'], r['step_kernel_ms'], r['apply_kernel_ms'], r['count_kernel_ms'], r['whole_step_frac'], j['e2e']['value']/1e9))
PY
done
done
