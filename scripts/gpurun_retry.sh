#!/bin/bash
# development aid: gpurun, retried while the pod answers "busy" (exit code 3); usage: scripts/gpurun_retry.sh TIMEOUT 'command'
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$1" -- "$2" > /root/repo/gpurun_out/.retry.log 2>&1
  rc=$?
  if ! grep -q "status=transient" /root/repo/gpurun_out/.retry.log; then break; fi
  sleep 45
done
tail -${3:-40} /root/repo/gpurun_out/.retry.log
