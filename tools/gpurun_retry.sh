#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout_s> <command...>   -- retries while the pod answers "transient/busy" (nothing charged)
log="$1"; shift; to="$1"; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient" "$log" || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
echo "gpurun rc=$rc attempt=$i" >> "$log"
