#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> <logfile> '<command>' [extra gpurun args]  — retries while the pod is busy (exit code 3)
t=$1; log=$2; cmd=$3; shift 3
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" --timeout $t -- "$cmd" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $log; then exit $rc; fi
  sleep 90
done
exit 3
