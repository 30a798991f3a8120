#!/bin/bash
# round 2, call 20: K > 200 top-K rounds (tests + timing), ncu --set full of the BPR W=1 step kernels
tag=${1:-r2M}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_topk_tensor.py tests/test_gpu_topk_metrics.py -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" ); tail -5 gpurun_out/${tag}_pytest.log
( timeout 600 python tools/topk_perf.py cml 200000 500000 128 1 1000 > gpurun_out/${tag}_topk1000.log 2>&1; echo "topk1000 rc=$?" ); tail -3 gpurun_out/${tag}_topk1000.log
( timeout 600 python tools/topk_perf.py cml 200000 500000 128 1 100 >> gpurun_out/${tag}_topk1000.log 2>&1; echo "topk100 rc=$?" ); tail -2 gpurun_out/${tag}_topk1000.log
CMD="python bench.py --workload c2-bpr --steps 4 --warmup 3 --no-cpu-baseline --no-other-configs --topk-users 0"
( timeout 600 $CMD > gpurun_out/${tag}_bpr_plain.json 2> gpurun_out/${tag}_bpr_plain.err; echo "plain rc=$?" )
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_step|k_apply_staged|k_count" --launch-skip 12 -c 3 -o gpurun_out/${tag}_bpr_full $CMD > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/${tag}_*
