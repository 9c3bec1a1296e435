#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout> <command...>   (retries while the pod answers busy: exit code 3)
LOG=$1; shift; TO=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1; rc=$?
  if grep -q "status=transient" $LOG || [ $rc -eq 3 ]; then sleep 60; continue; fi
  exit $rc
done
exit 3
