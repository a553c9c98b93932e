#!/bin/bash
# usage: scripts/gpurun_retry.sh <out-file> <gpurun args...>   -- retries while the pod answers "transient" (nothing is charged then)
out=$1; shift
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  if grep -q "status=transient" "$out"; then sleep 45; else exit 0; fi
done
