#!/bin/bash
# usage: [GPURUN_FLAGS='--gpus 8'] tools/gpurun_retry.sh <log> <timeout> <command string>   - retries while the pod answers busy (exit 3 / transient)
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $GPURUN_FLAGS --timeout $to -- "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient" "$log" || [ $rc -eq 3 ]; then sleep 120; continue; fi
  exit $rc
done
exit 3
