#!/usr/bin/env bash
# developer helper: gpurun with retries while the pod answers "transient" / busy (nothing charged)
# usage: scripts/gpurun_retry.sh <log> <timeout_s> <command...>
log=$1; shift; to=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1; rc=$?
  if grep -q "status=transient\|status=busy" "$log" || [ $rc -eq 3 ]; then sleep 150; continue; fi
  exit $rc
done
exit 3
