#!/bin/bash
# session 3, call D: deferred LN partial sums + skipped bias restaging: parity, role counters, encoder time
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_models_gpu.py tests/test_gemm_gpu.py -m gpu -x -q > gpurun_out/t_s3d.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_s3d.log
echo "== fc1"; TT_GEMM_DEBUG=4 timeout 120 python tools/ln_probe.py 384 1536 2 307200 3 2>&1 | grep "gemm dbg" | tail -2
echo "== qkv"; TT_GEMM_DEBUG=4 timeout 120 python tools/ln_probe.py 1536 1152 0 307200 3 2>&1 | grep "gemm dbg" | tail -2
timeout 300 python tools/dec_bench.py 2400 9600 2>&1 | grep "fused=1"
timeout 300 python tools/stage_bench.py 8 quick 2>&1 | tail -15
