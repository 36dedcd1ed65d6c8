#!/bin/bash
# usage: scripts/gpurun_retry.sh <log> <timeout> <gpus> <command...>   -- retries while the pod answers "busy" (rc 3)
log=$1; to=$2; gpus=$3; shift 3
for i in $(seq 1 12); do
  if [ "$gpus" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1; else /usr/local/graft/bin/gpurun --gpus $gpus --timeout $to -- "$@" > $log 2>&1; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
