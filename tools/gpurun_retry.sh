#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <timeout> [--gpus N] <command>   — retries while the pod is busy (nothing is charged)
log=$1; shift; to=$1; shift
extra=""
if [ "$1" == "--gpus" ]; then extra="--gpus $2"; shift; shift; fi
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun --timeout $to $extra -- "$@" > $log 2>&1
  if grep -q "status=transient" $log; then sleep 45; continue; fi
  break
done
