#!/bin/bash
# usage: tools/gpurun_retry_n.sh <ngpus> <timeout_s> <logfile> <command...>
N=$1; T=$2; LOG=$3; shift 3
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus $N --timeout $T -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc after $i attempt(s)" >> $LOG; exit $rc; fi
  sleep 150
done
echo "gave up" >> $LOG; exit 3
