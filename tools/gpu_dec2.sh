#!/bin/bash
mkdir -p gpurun_out
echo "== priorities on"; timeout 300 python tools/dec_bench.py 2400 9600 2>&1 | grep -v Warning | grep "fused=1"
echo "== priorities off"; TT_DEC_PRIO=0 timeout 300 python tools/dec_bench.py 2400 9600 2>&1 | grep -v Warning | grep "fused=1"
