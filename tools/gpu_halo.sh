#!/bin/bash
for shape in "8 1024 1024 64 64" "8 512 512 64 128" "8 512 512 64 32" "8 512 512 32 32"; do
  for halo in 1 0; do
    TT_CONV_HALO=$halo python tools/conv_probe.py $shape 5 2>&1 | grep -v Warn | sed "s/^/halo=$halo /"
    for dbg in 4 6; do TT_CONV_HALO=$halo TT_GEMM_DEBUG=$dbg python tools/conv_probe.py $shape 1 2>&1 | grep "gemm dbg" | tail -1 | sed "s/^/halo=$halo dbg=$dbg /"; done
  done
done
