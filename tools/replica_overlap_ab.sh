#!/bin/bash
# N-GPU A/B of the two-stream form of the 'replicate' transport: parity first, then the bench line with and without it
N=${1:-2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
$T 29521 tests/dist_check.py replicate 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -6
CF_DIST_BACKEND=gloo $T 29522 tests/dist_check.py replicate 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -4
B="bench.py --gpus $N --steps 40 --warmup 5 --no-c5 --no-other-configs --topk-users 0"
for v in 1 0; do
  echo "== CF_REPLICA_OVERLAP=$v"
  CF_REPLICA_OVERLAP=$v $T 2952$v $B 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('value %.4f G  ms %.4f  e2e %.4f G' % (d['value'] / 1e9, d['ms_per_step'], d['e2e']['value'] / 1e9))
        print(json.dumps(d['phases_ms_per_step']))
"
done
