#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> '<command>'   -- retries while the pod answers busy (exit 3) or transient
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun ${GPUS:+--gpus $GPUS} --timeout "$1" -- "$2" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" /tmp/gpurun_last.log; then break; fi
  sleep 90
done
tail -120 /tmp/gpurun_last.log
exit $rc
