#!/bin/bash
# gpu_retry_n.sh <gpus> <logfile> <timeout> <command...>
n=$1; log=$2; to=$3; shift 3
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --gpus $n --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "rc=$rc" >> $log; exit $rc; fi
  sleep 60
done
echo "gave up" >> $log
