#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_stages_gpu.py tests/test_e2e_gpu.py -m gpu -x -q -s -k "warp or rectify" > gpurun_out/t1.log 2>&1; echo "tests rc=$?"
grep -v "Warning\|warn\|^tests/\|key_padding\|^$" gpurun_out/t1.log | tail -30
