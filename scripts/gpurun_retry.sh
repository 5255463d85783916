#!/bin/bash
# gpurun_retry.sh <timeout> <command...>: retry while the pod answers busy (exit 3); log to gpurun_out/retry.log
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > /tmp/gpurun_last.txt 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then break; fi
  sleep 60
done
tail -40 /tmp/gpurun_last.txt
exit $rc
