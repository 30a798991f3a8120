#!/bin/bash
# round 2, call 28: cooperative-compaction low mark sweep + ncu --set full of k_topk_tc on the 500 k catalogue (one wave)
tag=${1:-r2U}
mkdir -p gpurun_out
for low in 160 200 250 300; do
echo "== CF_TC_COOP_LOW=$low" >> gpurun_out/${tag}_perf.log
( CF_TC_COOP_LOW=$low timeout 600 python tools/topk_perf.py cml 200000 500000 128 2 100 >> gpurun_out/${tag}_perf.log 2>&1 )
done
grep -v fallback gpurun_out/${tag}_perf.log
CMD="python tools/topk_perf.py cml 37888 500000 128 1 100"
( timeout 300 $CMD > gpurun_out/${tag}_plain.log 2>&1; echo "plain rc=$?" ); tail -1 gpurun_out/${tag}_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_topk_tc|k_rerank" --launch-skip 2 -c 2 -o gpurun_out/${tag}_topk500k $CMD > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
