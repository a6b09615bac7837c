#!/bin/bash
# scripts/gpurun_retry.sh TIMEOUT 'command' : gpurun, retried while the pod answers "busy" (exit code 3, nothing charged)
t=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 60
done
exit 3
