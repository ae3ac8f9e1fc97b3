#!/bin/bash
# usage: gpurun_retry.sh <max_tries> <gpurun args...>   (retries while the pod is busy)
tries=$1; shift
for i in $(seq 1 $tries); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | tail -60
  if ! echo "$out" | grep -q "status=transient"; then exit 0; fi
  sleep 75
done
exit 3
