#!/bin/bash
tag=${1:-run}
mkdir -p gpurun_out
for rep in a b c; do
( timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --topk-users 0 > gpurun_out/${tag}_bench_${rep}.json 2> gpurun_out/${tag}_bench_${rep}.err; echo "bench $rep rc=$?" )
python - <<PY
import json
j=json.loads(open('gpurun_out/${tag}_bench_${rep}.json').read().strip().splitlines()[-1])
print('$rep value %.3f G  ms %.3f  whole_step_frac %.3f e2e %.3f G' % (j['value']/1e9, j['ms_per_step'], j['roofline']['whole_step_frac'], j['e2e']['value']/1e9))
for k,v in j['other_configs'].items():
    print('   ', k, v.get('value'), v.get('ms_per_step', v.get('ms_per_half_sweep')), (v.get('roofline') or {}).get('whole_step_frac'))
PY
done
( timeout 600 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_steps.py tests/test_gpu_sampler.py tests/test_gpu_rating.py tests/test_gpu_tuples.py -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" ); tail -3 gpurun_out/${tag}_pytest.log
