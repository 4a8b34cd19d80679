#!/bin/bash
# session 3, call Q: proj folded into the fused MLP kernel: correctness probe, counters, timing A/B
PROBE_PROJ=1 timeout 120 python tools/mlp_probe.py 2 3 300 2>&1 | tail -4
TT_MLP_DEBUG=1 TT_ENC_PROJFUSE=1 timeout 200 python tools/dec_bench.py 2400 2>&1 | grep "mlp dbg" | tail -2
for v in 0 1 0 1; do echo "== TT_ENC_PROJFUSE=$v"; TT_ENC_PROJFUSE=$v timeout 200 python tools/dec_bench.py 2400 9600 2>&1 | grep "fused=1"; done
