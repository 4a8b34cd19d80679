#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_models_gpu.py tests/test_gemm_gpu.py -m gpu -x -q -k "parseq or layernorm_fused" > gpurun_out/t_ln2.log 2>&1; echo "tests rc=$?"
grep -v "Warning\|warn\|^tests/\|key_padding\|^$" gpurun_out/t_ln2.log | tail -12
echo "== LN fused"; timeout 300 python tools/stage_bench.py 8 lnf 2>&1 | grep -v Warning | tail -3; grep "^lin" gpurun_out/gemm_launches_lnf.csv | head -7
echo "== LN unfused"; TT_ENC_LNFUSE=0 timeout 300 python tools/stage_bench.py 8 lnu 2>&1 | grep -v Warning | tail -3; grep "^lin" gpurun_out/gemm_launches_lnu.csv | head -7
