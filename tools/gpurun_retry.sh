#!/bin/bash
# usage: tools/gpurun_retry.sh <gpurun args...>   -- retries while the pod answers "transient" (nothing charged)
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 150; continue; fi
  echo "$out" | tail -60
  exit 0
done
echo "gave up after 20 transient answers"
