#!/bin/bash
# ncu --set full of the specialised step kernels: BPR W=1 (configs[1] shape) and GBPR configs[2]
tag=${1:-r3C}
mkdir -p gpurun_out
for wl in c2-bpr c3; do
CMD="python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu-baseline --no-other-configs --topk-users 0"
( timeout 600 $CMD > gpurun_out/${tag}_${wl}_plain.json 2> gpurun_out/${tag}_${wl}_plain.err; echo "$wl plain rc=$?" )
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_step|k_apply_staged|k_count" --launch-skip 12 -c 3 -o gpurun_out/${tag}_${wl}_full $CMD > gpurun_out/${tag}_${wl}_ncu.log 2>&1; echo "$wl ncu rc=$?"
done
ls -la gpurun_out/${tag}_*
