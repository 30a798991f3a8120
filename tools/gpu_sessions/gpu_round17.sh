#!/bin/bash
tag=${1:-run}
mkdir -p gpurun_out
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-other-configs --topk-users 0"
( timeout 600 $CMD > gpurun_out/${tag}_plain.json 2> gpurun_out/${tag}_plain.err; echo "plain rc=$?" )
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_step|k_apply|k_count|k_sample|k_clip" -c 200 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu_launches.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_step|k_apply_staged|k_count" --launch-skip 12 -c 3 -o gpurun_out/${tag}_c2_full $CMD > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
grep -c k_step gpurun_out/${tag}_launches.csv
