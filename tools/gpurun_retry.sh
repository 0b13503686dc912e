#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun options] -- 'command'    (retries while the pod answers "busy"/transient)
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|nothing was charged"; then sleep 60; continue; fi
  echo "$out"; exit $rc
done
echo "gpurun_retry: gave up"; exit 3
