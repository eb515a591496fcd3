#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <timeout> '<command>' [extra gpurun args]  — retries while the pod answers "busy" (rc 3)
log=$1; to=$2; cmd=$3; shift 3
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" --timeout $to -- "$cmd" > $log 2>&1; rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 120
done
exit 3
