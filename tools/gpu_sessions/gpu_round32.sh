#!/bin/bash
# 2 GPUs: 'replicate' transport: parity (dist_check replicate / auto), pytest dist, bench N=2 (auto -> replicate on configs[1])
tag=${1:-r2Y}
mkdir -p gpurun_out
N=2
for tr in replicate auto; do
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/dist_check.py $tr > gpurun_out/${tag}_dist_check_N${N}_$tr.log 2>&1; echo "dist_check $tr rc=$?" ); grep -E "max\|diff|MISMATCH|False|DIST_CHECK" gpurun_out/${tag}_dist_check_N${N}_$tr.log | head -12
done
( timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_multi.py -x -q > gpurun_out/${tag}_pytest_dist.log 2>&1; echo "pytest dist rc=$?" ); tail -3 gpurun_out/${tag}_pytest_dist.log
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${tag}_bench_N${N}.json 2> gpurun_out/${tag}_bench_N${N}.err; echo "bench N$N rc=$?" )
python - <<PY
import json
j=json.loads(open('gpurun_out/${tag}_bench_N2.json').read().strip().splitlines()[-1])
print('N=2 value %.3f G ms %.3f e2e %.3f G' % (j['value']/1e9, j['ms_per_step'], j['e2e']['value']/1e9))
print('transport', j['item_transport'][:80])
print('phases', j.get('phases_ms_per_step'))
print('c5', (j.get('c5') or {}).get('value'), (j.get('c5') or {}).get('item_transport','')[:60])
print('topk', j['topk']['value'], j['topk']['item_sharded']['value'])
PY
tail -5 gpurun_out/${tag}_bench_N${N}.err
