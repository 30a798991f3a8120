#!/bin/bash
# round 2, call 27: cooperative compactions in the tensor top-K: tests + A/B
tag=${1:-r2T}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_topk_tensor.py tests/test_gpu_topk_metrics.py -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" ); tail -4 gpurun_out/${tag}_pytest.log
for c in 0 1; do
echo "== CF_TC_COOP=$c" >> gpurun_out/${tag}_perf.log
( CF_TC_COOP=$c timeout 600 python tools/topk_perf.py cml 200000 500000 128 2 100 >> gpurun_out/${tag}_perf.log 2>&1 )
( CF_TC_COOP=$c timeout 600 python tools/topk_perf.py bpr 200000 500000 128 2 10 >> gpurun_out/${tag}_perf.log 2>&1 )
( CF_TC_COOP=$c timeout 600 python tools/topk_perf.py cml 200000 500000 128 2 200 >> gpurun_out/${tag}_perf.log 2>&1 )
done
grep -v fallback gpurun_out/${tag}_perf.log
for c in 0 1; do
( CF_TC_COOP=$c timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs --topk-c5-items 0 > gpurun_out/${tag}_bench_topk_c$c.json 2> gpurun_out/${tag}_bench_topk_c$c.err; echo "bench rc=$?" )
python - <<PY
import json
j=json.loads(open('gpurun_out/${tag}_bench_topk_c$c.json').read().strip().splitlines()[-1])
t=j['topk']
print('coop=$c', t['value'], t['ms'], t['frac_of_tensor_peak'], t['candidates_per_row'])
PY
done
