#!/bin/bash
# 2-GPU session: the engine's own multi-GPU path (tests) + bench under torchrun with the engine_dp_check
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_e2e_gpu.py -m gpu -x -q -k "two_gpu or mixed or independent" > gpurun_out/t_n2.log 2>&1; echo "2-gpu tests rc=$?"; tail -3 gpurun_out/t_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
cut -c1-600 gpurun_out/bench_n2.json; grep -o '"engine_dp_check": "[a-z]*"' gpurun_out/bench_n2.json
