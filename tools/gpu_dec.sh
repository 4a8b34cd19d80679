#!/bin/bash
# fused decoder bring-up: parity tests of the PARSeq path, then stage timings
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_models_gpu.py -m gpu -x -q -s -k "parseq" > gpurun_out/t_dec.log 2>&1; echo "parseq tests rc=$?"
grep -v Warning gpurun_out/t_dec.log | tail -25
timeout 300 python tools/dec_bench.py 2>&1 | grep -v Warning | tail -8
timeout 300 python tools/latency_bench.py 2>&1 | grep -v Warning | tail -9
