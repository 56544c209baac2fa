#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> [--gpus N] -- '<command>'   (retries while the pod has no free slot)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  grep -q '"status": "transient"' gpurun_out/.last_call.json 2>/dev/null
  sleep 90
done
exit 3
