#!/bin/bash
# session 3, call S: proj + MLP kernel on by default: kernel tests, model tests, sanitizers, then the whole suite + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py -m gpu -x -q -s -k "fused_encoder_mlp" > gpurun_out/t_mlp.log 2>&1; echo "kernel test rc=$?"; grep -E "rel-L2|passed|failed|Error" gpurun_out/t_mlp.log | tail -10
timeout 600 compute-sanitizer --tool synccheck python -m pytest tests/test_gemm_gpu.py -m gpu -x -q -k "fused_encoder_mlp and 384" > gpurun_out/synccheck_mlp.log 2>&1; echo "synccheck rc=$?"; tail -2 gpurun_out/synccheck_mlp.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gemm_gpu.py -m gpu -x -q -k "fused_encoder_mlp and 384" > gpurun_out/memcheck_mlp.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/memcheck_mlp.log
bash tools/gpu_s3f.sh
