#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout> <command...>   - retries while the pod answers "transient" / busy (exit 3)
log=$1; to=$2; shift 2
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if grep -q "status=transient" $log || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
exit $rc
