#!/bin/bash
# role wait/busy counters of the GEMM kernel (TT_GEMM_DEBUG=4) for the K=384 shapes, 8 vs 16 epilogue warps
for ew in 8 16; do
for shape in "384 1536 2 0" "384 1152 0 0" "1536 384 0 1" "384 384 0 1"; do
  for dbg in 4 5 6; do
    echo "== EW=$ew shape=$shape debug=$dbg"
    TT_GEMM_EW=$ew TT_GEMM_DEBUG=$dbg timeout 120 python tools/gemm_probe.py $shape 0 0 2 2>&1 | grep "gemm dbg" | tail -1
  done
done
done
for ew in 8 16; do echo "== probe2 EW=$ew"; TT_GEMM_EW=$ew timeout 200 python tools/gemm_probe2.py 2>&1 | grep -v Warn; done
