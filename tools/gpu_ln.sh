#!/bin/bash
# LayerNorm-fusion bring-up: kernel pair test, PARSeq parity, stage timings fused vs unfused
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -x -q -s -k "layernorm_fused" > gpurun_out/t_ln.log 2>&1; echo "ln pair tests rc=$?"
grep -v Warning gpurun_out/t_ln.log | tail -12
timeout 600 python -m pytest tests/test_models_gpu.py -m gpu -x -q -s -k "parseq" > gpurun_out/t_ln2.log 2>&1; echo "parseq tests rc=$?"
grep -v "Warning\|warn\|^tests/\|key_padding\|^$" gpurun_out/t_ln2.log | tail -16
echo "== LN fused"; timeout 300 python tools/dec_bench.py 2400 9600 2>&1 | grep "fused=1"
echo "== LN unfused"; TT_ENC_LNFUSE=0 timeout 300 python tools/dec_bench.py 2400 9600 2>&1 | grep "fused=1"
