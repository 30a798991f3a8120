#!/bin/bash
# 2 GPUs: the driver's N=2 commands (ours + reference arm) after the bench_dist edits
tag=${1:-r3E}
mkdir -p gpurun_out
N=2
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${tag}_bench_N${N}.json 2> gpurun_out/${tag}_bench_N${N}.err; echo "bench N$N rc=$?" )
python - <<PY
import json
j=json.loads(open('gpurun_out/${tag}_bench_N2.json').read().strip().splitlines()[-1])
print('N=2 value %.3f G ms %.3f e2e %.3f G' % (j['value']/1e9, j['ms_per_step'], j['e2e']['value']/1e9))
print('config', json.dumps(j['config'])[:200])
print('sharding', j['sharding'][:200])
print('transport', j['item_transport'][:120])
print('c5', (j.get('c5') or {}).get('value'))
print('topk', j['topk']['value'], j['topk']['item_sharded']['value'])
PY
