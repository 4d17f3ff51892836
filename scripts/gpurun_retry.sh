#!/bin/bash
# usage: gpurun_retry.sh <timeout> <command...>  - retries while the pod answers busy (exit 3 / transient)
to=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  out=$(/usr/local/graft/bin/gpurun --timeout $to -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 60; continue; fi
  if [ $rc -eq 3 ]; then sleep 60; continue; fi
  echo "$out" | tail -40; exit $rc
done
echo "gave up: pod busy"; exit 3
