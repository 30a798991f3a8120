#!/bin/bash
# round 2, call 24 (2 GPUs): sharded step with the rest-of-chunk sampler on a third stream: parity (dist_check auto) + bench N=2; 2-GPU pytest
tag=${1:-r2Q}
mkdir -p gpurun_out
N=2
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/dist_check.py auto > gpurun_out/${tag}_dist_check_N${N}_auto.log 2>&1; echo "dist_check auto rc=$?" ); tail -4 gpurun_out/${tag}_dist_check_N${N}_auto.log
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${tag}_bench_N${N}.json 2> gpurun_out/${tag}_bench_N${N}.err; echo "bench N$N rc=$?" )
python - <<PY
import json
j=json.loads(open('gpurun_out/${tag}_bench_N2.json').read().strip().splitlines()[-1])
print('N=2 value %.3f G ms %.3f e2e %.3f G' % (j['value']/1e9, j['ms_per_step'], j['e2e']['value']/1e9))
print('phases', j.get('phases_ms_per_step'))
c5=j.get('c5') or (j.get('other_configs') or {}).get('c5')
print('c5', (c5 or {}).get('value'))
print('topk', {k:(v.get('value') if isinstance(v,dict) else v) for k,v in (j.get('topk') or {}).items() if k in ('value','item_sharded','c5_items')})
PY
( timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_gpu_multi.py -x -q > gpurun_out/${tag}_pytest_dist.log 2>&1; echo "pytest dist rc=$?" ); tail -3 gpurun_out/${tag}_pytest_dist.log
