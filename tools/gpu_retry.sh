#!/bin/bash
# gpu_retry.sh <logfile> <timeout> <command...> : gpurun with retries while the pod answers "busy" (exit code 3)
log=$1; to=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "rc=$rc" >> $log; exit $rc; fi
  sleep 45
done
echo "gave up" >> $log
