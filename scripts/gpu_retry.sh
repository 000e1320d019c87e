#!/bin/bash
# usage: gpu_retry.sh <logfile> <timeout-seconds> <command...>   -- retries while the pod answers "busy" (exit 3)
LOG=$1; shift; TO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
