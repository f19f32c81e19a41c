#!/bin/bash
# retry a gpurun call while the pod answers "busy" (exit 3 / transient: nothing charged); usage: gpurun_retry.sh <timeout> <command...>
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if ! grep -q "status=transient" /tmp/gpurun_last.log; then cat /tmp/gpurun_last.log; exit $rc; fi
  sleep 60
done
cat /tmp/gpurun_last.log; exit 3
