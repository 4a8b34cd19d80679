#!/bin/bash
# GPU session: pair-mode GEMM correctness, then schedule / chunk sweeps (development aid)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q > gpurun_out/t_gemm.log 2>&1; echo "gemm tests rc=$?"
tail -5 gpurun_out/t_gemm.log
for cfg in "0 0 base" "1 0 pair" "1 148 pair_c148" "1 296 pair_c296" "0 148 single_c148" "1 74 pair_c74"; do
  set -- $cfg
  TT_GEMM_PAIR=$1 TT_ENC_CHUNK=$2 timeout 300 python tools/stage_bench.py 8 $3 2>&1 | grep -v Warning | tail -4
done
