#!/bin/bash
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q 2>&1 | tail -5
echo "== TE on"; timeout 200 python tools/gemm_probe3.py 2>&1 | grep -E "f321" 
echo "== TE off"; TT_GEMM_TE=0 timeout 200 python tools/gemm_probe3.py 2>&1 | grep -E "f321"
for shape in "1536 384 0 1" "384 384 0 1"; do for dbg in 4 5 6; do echo "== TE shape=$shape debug=$dbg"; TT_GEMM_DEBUG=$dbg timeout 120 python tools/gemm_probe.py $shape 0 0 2 2>&1 | grep "gemm dbg" | tail -1; done; done
