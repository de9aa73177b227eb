#!/bin/bash
# Development aid: submit a gpurun job, retrying while the pod answers "busy" (exit code 3).
#   tools/gpu.sh TIMEOUT 'command'  -> log in /tmp/gpu_last.log
t=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@" > /tmp/gpu_last.log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" /tmp/gpu_last.log; then exit $rc; fi
  sleep 60
done
exit 3
