#!/bin/bash
# session 3, call V: sanitizer evidence for the kernels added in this session
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gemm_gpu.py -m gpu -x -q -k "fused_encoder_mlp and 384" > gpurun_out/racecheck_mlp.log 2>&1; echo "racecheck mlp rc=$?"; tail -3 gpurun_out/racecheck_mlp.log
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize_probe.py > gpurun_out/memcheck_probe.log 2>&1; echo "memcheck probe rc=$?"; tail -3 gpurun_out/memcheck_probe.log
timeout 300 compute-sanitizer --tool synccheck python tools/sanitize_probe.py > gpurun_out/synccheck_probe.log 2>&1; echo "synccheck probe rc=$?"; tail -3 gpurun_out/synccheck_probe.log
