#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_models_gpu.py -m gpu -x -q -k "parseq" > gpurun_out/t_dec.log 2>&1; echo "parseq tests rc=$?"; tail -2 gpurun_out/t_dec.log
echo "== carveout max-shared"; timeout 300 python tools/dec_bench.py 300 2400 9600 2>&1 | grep "fused=1"
echo "== carveout default"; TT_DEC_CARVEOUT=0 timeout 300 python tools/dec_bench.py 300 2400 9600 2>&1 | grep "fused=1"
TT_DEC_DEBUG=1 python tools/dec_bench.py 9600 2>&1 | grep "dec dbg" | grep "mode 0" | head -2
TT_DEC_DEBUG=1 python tools/dec_bench.py 9600 2>&1 | grep "dec dbg" | grep "mode 1" | head -2
