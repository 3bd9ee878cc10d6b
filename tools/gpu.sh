#!/bin/bash
# usage: tools/gpu.sh <logfile> [--gpus N] <timeout> <command...>   -- gpurun with retries while the pod is busy
LOG=$1; shift
GP=""
if [ "$1" = "--gpus" ]; then GP="--gpus $2"; shift 2; fi
TO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $GP --timeout $TO -- "$@" > $LOG 2>&1
  if grep -q "status=transient" $LOG || grep -q "rc=3" $LOG; then sleep 45; continue; fi
  break
done
