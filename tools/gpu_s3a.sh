#!/bin/bash
# session 3, call A: validate the 64-column-chunk MLP pipelining of k_dec_dense<B>
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_models_gpu.py -m gpu -x -q -k "parseq" > gpurun_out/t_dec.log 2>&1; echo "parseq tests rc=$?"; tail -3 gpurun_out/t_dec.log
timeout 300 python tools/dec_bench.py 300 2400 9600 2>&1 | grep "fused=" 
TT_DEC_DEBUG=1 timeout 120 python tools/dec_bench.py 9600 2>&1 | grep "dec dbg" | grep "mode 0" | head -2
TT_DEC_DEBUG=1 timeout 120 python tools/dec_bench.py 9600 2>&1 | grep "dec dbg" | grep "mode 1" | head -2
