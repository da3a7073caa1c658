#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <gpurun args...>   - retries while the pod answers "busy" (exit 3), up to 12 times
log=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 120
done
exit 3
