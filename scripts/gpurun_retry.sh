#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout-seconds> <gpus> <command...>   (retries while the pod answers "busy", exit code 3)
t=$1; shift
n=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $t --gpus $n -- "$@"
  rc=$?
  [ $rc -ne 3 ] && exit $rc
  sleep 90
done
exit 3
