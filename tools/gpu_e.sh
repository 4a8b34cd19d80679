#!/bin/bash
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | grep -E "Error|passed|failed" | cut -c1-600
for shape in "8 1024 1024 64 64" "8 512 512 64 128" "8 512 512 64 32" "8 512 512 32 32" "8 256 256 128 64"; do
  for sp in 1 0; do
    TT_CONV_SPLIT2=$sp python tools/conv_probe.py $shape 5 2>&1 | grep -v Warn | sed "s/^/split2=$sp /"
  done
done
