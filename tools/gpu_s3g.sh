#!/bin/bash
# session 3, call G: per-crop early exit of the AR loop: parity + decoder time
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_models_gpu.py -m gpu -x -q -s -k "parseq" > gpurun_out/t_ee.log 2>&1; echo "parseq tests rc=$?"; grep -E "mean AR steps|passed|failed|Error|error" gpurun_out/t_ee.log | tail -12
for v in 1 0; do echo "== TT_DEC_EARLY_EXIT=$v"; TT_DEC_EARLY_EXIT=$v timeout 300 python tools/dec_bench.py 300 2400 9600 2>&1 | grep "fused=1"; done
timeout 900 python -m pytest tests/test_e2e_gpu.py tests/test_fixture_images.py -m gpu -x -q > gpurun_out/t_e2e.log 2>&1; echo "e2e tests rc=$?"; tail -2 gpurun_out/t_e2e.log
