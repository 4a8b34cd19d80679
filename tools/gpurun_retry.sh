#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <command...>   -- retries while the pod answers busy (exit 3), nothing is charged for those
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > /tmp/gpurun_last.txt 2>&1; rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.txt; then echo "[retry $i] busy"; sleep 150; continue; fi
  cat /tmp/gpurun_last.txt; exit $rc
done
echo "gave up"; exit 3
