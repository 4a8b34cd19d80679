#!/bin/bash
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | grep -E "Error|passed|failed" | cut -c1-600
for shape in "8 1024 1024 64 64" "8 512 512 64 128" "8 512 512 32 32"; do
  for halo in 1 0; do
    TT_CONV_HALO=$halo python tools/conv_probe.py $shape 5 2>&1 | grep -v Warn | sed "s/^/halo=$halo /"
  done
done
TT_GEMM_DEBUG=4 python tools/conv_probe.py 8 1024 1024 64 64 1 2>&1 | grep "gemm dbg" | tail -1
timeout 200 python tools/gemm_probe3.py 2>&1 | grep -E "BN192|BN256"
