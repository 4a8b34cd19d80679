#!/bin/bash
# session 3, call K (2 GPUs): the engine's own multi-GPU tests + the 2-rank bench line with engine_dp_check
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "two_gpu or mixed_page_sizes or independent_units" > gpurun_out/t_n2.log 2>&1; echo "2-GPU tests rc=$?"; tail -3 gpurun_out/t_n2.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; head -c 400 gpurun_out/bench_n2.json; echo
