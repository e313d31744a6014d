#!/bin/bash
# usage: tools/gpurun_retry.sh LOGFILE TIMEOUT 'command' [gpus]   -- retries while the pod answers "busy" (rc 3)
log=$1; to=$2; cmd=$3; gpus=${4:-1}
for i in $(seq 1 40); do
  if [ "$gpus" = "1" ]; then
    /usr/local/graft/bin/gpurun --timeout $to -- "$cmd" > $log 2>&1
  else
    /usr/local/graft/bin/gpurun --gpus $gpus --timeout $to -- "$cmd" > $log 2>&1
  fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
