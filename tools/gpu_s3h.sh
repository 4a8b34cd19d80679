#!/bin/bash
# session 3, call H: default bench line with the early-exit decoder
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; head -c 300 gpurun_out/bench.json; echo
tail -3 gpurun_out/bench.err
