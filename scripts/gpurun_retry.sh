#!/bin/bash
# usage: scripts/gpurun_retry.sh [gpurun options] -- 'command'   (retries while the pod answers "transient"/busy)
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | tail -40
  if echo "$out" | grep -q "status=transient\|rc=3\|no box"; then sleep 60; continue; fi
  break
done
