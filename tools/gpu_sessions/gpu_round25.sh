#!/bin/bash
# round 2, call 25: tensor top-K threshold warm-up: parity tests + A/B timing
tag=${1:-r2R}
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_topk_tensor.py tests/test_gpu_topk_metrics.py -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" ); tail -4 gpurun_out/${tag}_pytest.log
for w in 0 128 32; do
echo "== CF_TC_WARM=$w" >> gpurun_out/${tag}_perf.log
( CF_TC_WARM=$w timeout 600 python tools/topk_perf.py cml 200000 500000 128 2 100 >> gpurun_out/${tag}_perf.log 2>&1 )
( CF_TC_WARM=$w timeout 600 python tools/topk_perf.py bpr 200000 500000 128 2 10 >> gpurun_out/${tag}_perf.log 2>&1 )
( CF_TC_WARM=$w timeout 600 python tools/topk_perf.py cml 37888 10000000 128 1 100 >> gpurun_out/${tag}_perf.log 2>&1 )
done
( timeout 600 python tools/topk_perf.py cml 200000 500000 128 1 1000 >> gpurun_out/${tag}_perf.log 2>&1 )
( timeout 600 python tools/topk_perf.py gbpr 200000 27000 64 2 100 >> gpurun_out/${tag}_perf.log 2>&1 )
( CF_TC_WARM=0 timeout 600 python tools/topk_perf.py gbpr 200000 27000 64 2 100 >> gpurun_out/${tag}_perf.log 2>&1 )
grep -v fallback gpurun_out/${tag}_perf.log
