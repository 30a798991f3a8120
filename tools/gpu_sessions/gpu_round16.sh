#!/bin/bash
tag=${1:-run}
mkdir -p gpurun_out
( timeout 180 python -m pytest tests/test_gpu_topk_tensor.py -x -q -k paired > gpurun_out/${tag}_pytest_pair.log 2>&1; echo "pytest pair rc=$?" ); tail -15 gpurun_out/${tag}_pytest_pair.log
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
( timeout 300 python -m pytest tests/test_gpu_topk_tensor.py tests/test_gpu_topk_metrics.py tests/test_gpu_full_size.py -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" ); tail -4 gpurun_out/${tag}_pytest.log
for pr in 1 0; do
( CF_TC_PAIR=$pr timeout 300 python tools/topk_perf.py bpr 1000000 10000000 128 1 > gpurun_out/${tag}_perf_pair${pr}.log 2>&1; echo "perf pair=$pr rc=$?" ); tail -2 gpurun_out/${tag}_perf_pair${pr}.log
done
