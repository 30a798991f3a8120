#!/bin/bash
# round 2, call 26: re-rank with dynamic shared memory / 128-thread blocks + warm-up: tests, perf, bench top-K objects
tag=${1:-r2S}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_topk_tensor.py tests/test_gpu_topk_metrics.py tests/test_gpu_full_size.py -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" ); tail -4 gpurun_out/${tag}_pytest.log
( timeout 600 python tools/topk_perf.py cml 200000 500000 128 2 100 >> gpurun_out/${tag}_perf.log 2>&1 )
( timeout 600 python tools/topk_perf.py bpr 200000 500000 128 2 10 >> gpurun_out/${tag}_perf.log 2>&1 )
( timeout 600 python tools/topk_perf.py cml 200000 500000 128 1 1000 >> gpurun_out/${tag}_perf.log 2>&1 )
grep -v fallback gpurun_out/${tag}_perf.log
( timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/${tag}_bench_topk.json 2> gpurun_out/${tag}_bench_topk.err; echo "bench rc=$?" )
python - <<PY
import json
j=json.loads(open('gpurun_out/${tag}_bench_topk.json').read().strip().splitlines()[-1])
t=j['topk']
print(json.dumps({k:v for k,v in t.items() if not isinstance(v,(dict,list))}, indent=0)[:1500])
for k,v in t.items():
    if isinstance(v,dict): print(k, json.dumps({a:b for a,b in v.items() if not isinstance(b,(dict,list))})[:900])
PY
