#!/bin/bash
# session 3, call E: same-box A/B of the epilogue preamble change (TT_GEMM_DEBUG=8 = legacy behaviour)
for v in 0 8 0 8; do echo "== TT_GEMM_DEBUG=$v"; TT_GEMM_DEBUG=$v timeout 300 python tools/dec_bench.py 9600 2>&1 | grep "fused=1"; done
for v in 0 8; do echo "== TT_GEMM_DEBUG=$v"; TT_GEMM_DEBUG=$v timeout 300 python tools/stage_bench.py 8 quick 2>&1 | tail -2; cp gpurun_out/gemm_launches_quick.csv gpurun_out/gemm_launches_dbg$v.csv; done
