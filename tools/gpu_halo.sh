#!/bin/bash
for v in 1 2; do echo "== TT_CONV_HALO=$v"; TT_CONV_HALO=$v timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q -k "halo" 2>&1 | grep -E "Error|passed|failed" | cut -c1-500; done
