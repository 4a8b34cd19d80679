#!/bin/bash
# GPU session: parity suite + per-stage timings + default bench (development aid)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?"
tail -4 gpurun_out/t_gpu.log
timeout 300 python tools/stage_bench.py 8 full 2>&1 | grep -v Warning | tail -8
if [ "$1" != "nobench" ]; then
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench.json
fi
