#!/bin/bash
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | grep -E "Error|passed|failed" | cut -c1-600
echo "== TS on"; timeout 200 python tools/gemm_probe3.py 2>&1 | grep -E "f320" 
echo "== TS off"; TT_GEMM_TS=0 timeout 200 python tools/gemm_probe3.py 2>&1 | grep -E "f320"
echo "== TS on, EW16 everywhere"; TT_GEMM_EW=16 timeout 200 python tools/gemm_probe3.py 2>&1 | grep -E "f320"
for dbg in 4 5 6; do echo "== TS fc1 debug=$dbg"; TT_GEMM_DEBUG=$dbg timeout 120 python tools/gemm_probe.py 384 1536 2 0 0 0 2 2>&1 | grep "gemm dbg" | tail -1; done
