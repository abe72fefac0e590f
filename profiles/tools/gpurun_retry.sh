#!/bin/bash
# usage: gpurun_retry.sh <log> <gpurun args...>   retry while the pod answers "busy" (exit code 3 / status=transient)
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > $LOG 2>&1
  if grep -q "status=transient\|rc=3\|answers busy" $LOG && ! grep -q "status=ok" $LOG; then sleep 90; continue; fi
  break
done
