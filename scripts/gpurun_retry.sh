#!/bin/bash
# usage: scripts/gpurun_retry.sh <log> <gpurun args...>   -- retries while gpurun answers 3 (no box / slot free, nothing charged)
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
