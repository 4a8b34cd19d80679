#!/bin/bash
# session 3, call C: role counters of the LayerNorm-consumer GEMMs (fc1-like: N 1536 GELU; qkv-like: N 1152)
for ew in 16 8; do
echo "== fc1 (proj producer + fc1 consumer), EW=$ew"; TT_GEMM_EW=$ew TT_GEMM_DEBUG=4 timeout 120 python tools/ln_probe.py 384 1536 2 307200 3 2>&1 | grep "gemm dbg" | tail -2
echo "== qkv (fc2 producer + qkv consumer), EW=$ew"; TT_GEMM_EW=$ew TT_GEMM_DEBUG=4 timeout 120 python tools/ln_probe.py 1536 1152 0 307200 3 2>&1 | grep "gemm dbg" | tail -2
done
