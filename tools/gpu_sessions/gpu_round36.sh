#!/bin/bash
# GBPR specialised kernel with 128-thread blocks (5 blocks = 20 warps per SM) vs 256-thread blocks: parity tests + A/B on configs[2]
tag=${1:-r3D}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_steps.py -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" ); tail -3 gpurun_out/${tag}_pytest.log
run() {
( env $1 timeout 600 python bench.py --workload c3 --steps 30 --warmup 5 --no-cpu-baseline --no-other-configs --topk-users 0 > gpurun_out/${tag}_c3_$2.json 2> gpurun_out/${tag}_c3_$2.err; echo "bench c3 $2 rc=$?" )
python - <<PY
import json
j=json.loads(open('gpurun_out/${tag}_c3_$2.json').read().strip().splitlines()[-1])
r=j['roofline']
print('c3 $2 value %.3f G  ms %.3f  step %.3f apply %.3f count %.3f  e2e %.3f G' % (j['value']/1e9, j['ms_per_step'], r['step_kernel_ms'], r['apply_kernel_ms'], r['count_kernel_ms'], j['e2e']['value']/1e9))
PY
}
run CF_STEP_FAST_T256=1 t256
run CF_STEP_FAST_T256=0 t128
