#!/bin/bash
# usage: scripts/gpurun_retry.sh <logfile> <gpurun args...>   — retries while the pod answers "busy" (exit 3), up to 40 times
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc after $i attempt(s)" >> "$log"; exit $rc; fi
  sleep 100
done
echo "gave up after 40 busy answers" >> "$log"; exit 3
