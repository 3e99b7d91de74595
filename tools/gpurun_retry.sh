#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...> : retries while the pod answers "transient / busy" (nothing charged)
log=$1; shift
for i in $(seq 1 20); do
  gpurun "$@" > "$log" 2>&1
  if grep -q "status=transient\|status=busy\|rc=3" "$log" || grep -q "retry in a few minutes" "$log"; then sleep 120; continue; fi
  break
done
