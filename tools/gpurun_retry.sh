#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <logfile> <command...>   -- retries while the pod answers busy (exit 3 / transient)
TO=$1; LOG=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" $LOG || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
echo "gpurun_retry done rc=$rc" >> $LOG
