#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout> <command...>  — retries while the pod answers "busy" (exit 3), up to ~40 min
LOG=$1; TMO=$2; shift 2
for i in $(seq 1 14); do
  /usr/local/graft/bin/gpurun --timeout $TMO -- "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $LOG; then exit $rc; fi
  sleep 60
done
exit 3
