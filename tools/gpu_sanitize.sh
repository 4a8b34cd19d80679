#!/bin/bash
# compute-sanitizer over the round-2 kernels (tools/sanitize_probe.py): memcheck, synccheck, racecheck
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_probe.py > gpurun_out/$tool.log 2>&1; echo "$tool rc=$?"
  grep -v Warn gpurun_out/$tool.log | tail -3
done
