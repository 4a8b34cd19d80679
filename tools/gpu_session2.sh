#!/bin/bash
# GPU session: full parity suite on HEAD, pair/no-pair sweeps, default bench (development aid)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?"
tail -5 gpurun_out/t_gpu.log
for cfg in "0 0 base" "1 0 pair"; do
  set -- $cfg
  TT_GEMM_PAIR=$1 TT_ENC_CHUNK=$2 timeout 300 python tools/stage_bench.py 8 $3 2>&1 | grep -v Warning | tail -4
done
timeout 300 python tools/stage_bench.py 8 full 2>&1 | grep -v Warning | tail -8
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json
