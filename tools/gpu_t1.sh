#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_models_gpu.py -m gpu -x -q -s -k "craft" > gpurun_out/t1.log 2>&1; echo "tests rc=$?"
grep -v "Warning\|warn\|^tests/\|key_padding\|^$" gpurun_out/t1.log | tail -14
echo "== pool fused"; timeout 300 python tools/stage_bench.py 8 pf 2>&1 | grep -v Warning | tail -2
echo "== pool unfused"; TT_CRAFT_POOLFUSE=0 timeout 300 python tools/stage_bench.py 8 pu 2>&1 | grep -v Warning | tail -2
grep "^conv" gpurun_out/gemm_launches_pf.csv | head -12 | cut -d, -f1,3
echo; grep "^conv" gpurun_out/gemm_launches_pu.csv | head -12 | cut -d, -f1,3
