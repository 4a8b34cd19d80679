#!/bin/bash
# session 3, call P: fused encoder MLP on by default: kernel test, model tests, sanitizer (synccheck) on the kernel probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py -m gpu -x -q -s -k "fused_encoder_mlp" > gpurun_out/t_mlp.log 2>&1; echo "kernel test rc=$?"; grep -E "rel-L2|passed|failed|Error" gpurun_out/t_mlp.log | tail -8
timeout 1200 python -m pytest tests/test_models_gpu.py -m gpu -x -q -s -k "parseq" > gpurun_out/t_mlp2.log 2>&1; echo "parseq tests rc=$?"; grep -E "fused-MLP|passed|failed|Error" gpurun_out/t_mlp2.log | tail -6
timeout 600 compute-sanitizer --tool synccheck python -m pytest tests/test_gemm_gpu.py -m gpu -x -q -k "fused_encoder_mlp and 384" > gpurun_out/synccheck_mlp.log 2>&1; echo "synccheck rc=$?"; tail -4 gpurun_out/synccheck_mlp.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gemm_gpu.py -m gpu -x -q -k "fused_encoder_mlp and 384" > gpurun_out/memcheck_mlp.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck_mlp.log
