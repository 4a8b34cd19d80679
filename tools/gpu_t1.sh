#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_models_gpu.py -m gpu -x -q -s -k "craft" > gpurun_out/t1.log 2>&1; echo "tests rc=$?"
grep -v "Warning\|warn\|^tests/\|key_padding\|^$" gpurun_out/t1.log | tail -14
echo "== c1_1 from u8"; timeout 300 python tools/stage_bench.py 8 cn 2>&1 | grep -v Warning | tail -2
echo "== c1_1 gemm path"; TT_CRAFT_C11=0 timeout 300 python tools/stage_bench.py 8 co 2>&1 | grep -v Warning | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_conv1_1|k_im2col" -c 4 --csv python tools/stage_bench.py 8 quick 2>/dev/null | grep -E "k_conv1_1|k_im2col" | awk -F, '{print $5, $NF}' | head -4
