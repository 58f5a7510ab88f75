#!/bin/bash
# gpurun with retries while the pod answers busy (exit code 3: nothing charged)
# usage: scripts/gpurun_retry.sh <log> <timeout-seconds> <command...>
log=$1; shift; to=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1; rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
