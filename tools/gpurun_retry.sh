#!/bin/bash
# usage: tools/gpurun_retry.sh <tag> <timeout> <command...>   -- retries while the pod answers busy (exit 3)
TAG=$1; shift; TO=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > gpurun_out/${TAG}_gpurun.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
