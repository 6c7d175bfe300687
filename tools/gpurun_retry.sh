#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>' [extra gpurun flags...]
# Retries while the pod answers "busy / no slot" (exit 3); anything else is final.
T=$1; CMD=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" --timeout "$T" -- "$CMD"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
