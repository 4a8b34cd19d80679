#!/bin/bash
# session 3, call F: the whole GPU suite + the default bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_gpu.log
timeout 1500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; head -c 600 gpurun_out/bench.json; echo
python __graft_entry__.py smoke 2>&1 | tail -2
