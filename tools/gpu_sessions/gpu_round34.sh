#!/bin/bash
# 8 GPUs: 'replicate' transport (reduce-scatter + shard apply + all-gather form): parity + bench of configs[1] (auto)
tag=${1:-r3B}
mkdir -p gpurun_out
N=${2:-8}
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/dist_check.py replicate > gpurun_out/${tag}_dist_check_N${N}_replicate.log 2>&1; echo "dist_check replicate rc=$?" ); grep -E "max\|diff|MISMATCH|False|DIST_CHECK" gpurun_out/${tag}_dist_check_N${N}_replicate.log | head -12
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 20 --warmup 5 --no-c5 --topk-c5-items 0 > gpurun_out/${tag}_bench_N${N}.json 2> gpurun_out/${tag}_bench_N${N}.err; echo "bench N$N rc=$?" )
python - <<PY
import json
j=json.loads(open('gpurun_out/${tag}_bench_N${N}.json').read().strip().splitlines()[-1])
print('N=$N value %.3f G ms %.3f e2e %.3f G' % (j['value']/1e9, j['ms_per_step'], j['e2e']['value']/1e9))
print('transport', j['item_transport'][:60])
print('phases', {k[:34]: round(v,3) for k,v in (j.get('phases_ms_per_step') or {}).items()})
t=j.get('topk')
if t: print('topk', t['value'], t['item_sharded']['value'])
PY
