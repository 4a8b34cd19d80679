#!/bin/bash
# session 3, call L: fused encoder MLP kernel: correctness probe, then timing
timeout 120 python tools/mlp_probe.py 2 3 300 2>&1 | tail -6
for v in 0 1; do echo "== TT_ENC_MLPFUSE=$v"; TT_ENC_MLPFUSE=$v timeout 200 python tools/dec_bench.py 2400 9600 2>&1 | grep "fused=1"; done
