#!/bin/bash
# round 2, call 23: specialised step kernels incl. GBPR and the 16-lane form: parity tests + A/B timing of configs[2]
tag=${1:-r2P}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_steps.py tests/test_gpu_e2e.py tests/test_gpu_full_size.py -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" ); tail -5 gpurun_out/${tag}_pytest.log
run() {
( env $1 timeout 600 python bench.py --workload c3 --steps 30 --warmup 5 --no-cpu-baseline --no-other-configs --topk-users 0 > gpurun_out/${tag}_c3_$2.json 2> gpurun_out/${tag}_c3_$2.err; echo "bench c3 $2 rc=$?" )
python - <<PY
import json
j=json.loads(open('gpurun_out/${tag}_c3_$2.json').read().strip().splitlines()[-1])
r=j['roofline']
print('c3 $2 value %.3f G  ms %.3f  step %.3f apply %.3f count %.3f  whole_step_frac %.3f e2e %.3f G' % (j['value']/1e9, j['ms_per_step'], r['step_kernel_ms'], r['apply_kernel_ms'], r['count_kernel_ms'], r['whole_step_frac'], j['e2e']['value']/1e9))
PY
}
run CF_STEP_GENERIC=1 generic
run CF_STEP_FAST_NBUF=1 nbuf1
run CF_STEP_FAST_NBUF=2 nbuf2
