#!/bin/bash
# usage: gpurun_retry.sh <timeout> <out-file> <command...>   -- retries while the pod answers busy (nothing is charged)
t=$1; out=$2; shift 2
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout "$t" $GPURUN_FLAGS -- "$@" > "$out" 2>&1
  if grep -q "status=transient\|status=busy\|rc=None" "$out"; then sleep 120; else break; fi
done
