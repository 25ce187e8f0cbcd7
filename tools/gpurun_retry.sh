#!/bin/bash
# usage: tools/gpurun_retry.sh LOGFILE [gpurun args...]   -- retries while the pod answers busy (exit code 3)
LOG=$1; shift
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$LOG"; then exit $rc; fi
  sleep 90
done
exit 3
