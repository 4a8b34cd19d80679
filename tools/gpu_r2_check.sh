#!/bin/bash
# Round-2 GPU session: parity suite + per-stage timings + default bench (development aid)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?"
tail -6 gpurun_out/t_gpu.log
timeout 300 python tools/stage_bench.py 8 full 2>&1 | grep -v Warning | tail -12
timeout 300 python tools/latency_bench.py 2>&1 | grep -v Warning | tail -8
if [ "$1" != "nobench" ]; then
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
fi
