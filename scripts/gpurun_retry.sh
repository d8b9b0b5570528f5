#!/bin/bash
# usage: scripts/gpurun_retry.sh <tag> <timeout> <command...>   (retries while the pod answers "transient"/busy)
TAG=$1; shift; TMO=$1; shift
for attempt in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout $TMO -- "$@" > gpurun_out/${TAG}_call.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
