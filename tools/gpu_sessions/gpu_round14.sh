#!/bin/bash
# N GPUs: parity script + the driver's N>1 bench launch
tag=${1:-run}; N=${2:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
( timeout 600 $TR tests/dist_check.py auto > gpurun_out/${tag}_dist_check_N${N}_auto.log 2>&1; echo "dist_check auto rc=$?" ); tail -4 gpurun_out/${tag}_dist_check_N${N}_auto.log
t0=$(date +%s)
( timeout 1500 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${tag}_bench_N${N}.json 2> gpurun_out/${tag}_bench_N${N}.err; echo "bench N=$N rc=$? $(( $(date +%s) - t0 )) s" )
tail -3 gpurun_out/${tag}_bench_N${N}.err
python - <<PY
import json
try:
    j=json.loads(open('gpurun_out/${tag}_bench_N${N}.json').read().strip().splitlines()[-1])
    print('value %.3f G  ms %.3f  e2e %.3f G' % (j['value']/1e9, j['ms_per_step'], j['e2e']['value']/1e9))
    print('phases', j.get('phases_ms_per_step'))
    t=j.get('topk') or {}
    print('topk user-sharded %.0f users/s (%.3f per GPU)  item-sharded %.0f' % (t.get('value',0), t.get('frac_of_tensor_peak_per_gpu',0), (t.get('item_sharded') or {}).get('value',0)))
    c=j.get('c5') or {}
    print('c5 value %.3f G ms %.3f' % (c.get('value',0)/1e9, c.get('ms_per_step',0)))
    t=c.get('topk') or {}
    print('c5 topk user-sharded %.0f users/s (%.3f per GPU)  item-sharded %.0f' % (t.get('value',0), t.get('frac_of_tensor_peak_per_gpu',0), (t.get('item_sharded') or {}).get('value',0)))
except Exception as e: print('failed', e)
PY
