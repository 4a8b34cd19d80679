#!/bin/bash
# session 3, call B: one-pass softmax in k_attn_enc (A/B) + parity
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_models_gpu.py -m gpu -x -q -k "parseq" > gpurun_out/t_attn.log 2>&1; echo "parseq tests rc=$?"; tail -3 gpurun_out/t_attn.log
for v in 1 0 1 0; do echo "== TT_ATTN_ONEPASS=$v"; TT_ATTN_ONEPASS=$v timeout 300 python tools/dec_bench.py 2400 9600 2>&1 | grep "fused=1"; done
