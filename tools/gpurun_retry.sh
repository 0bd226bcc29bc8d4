#!/bin/bash
# usage: tools/gpurun_retry.sh [--gpus N] TIMEOUT SCRIPT...   -- retries while the pod answers "transient" (nothing charged)
GP=""
if [ "$1" == "--gpus" ]; then GP="--gpus $2"; shift 2; fi
TO=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun $GP --timeout $TO -- "$@" > /tmp/gpurun_last.log 2>&1
  if grep -q "status=transient" /tmp/gpurun_last.log; then sleep 90; continue; fi
  break
done
tail -60 /tmp/gpurun_last.log
