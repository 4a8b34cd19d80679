#!/bin/bash
# session 3, call M: role counters of the fused encoder MLP kernel
TT_MLP_DEBUG=1 TT_ENC_MLPFUSE=1 timeout 200 python tools/dec_bench.py 2400 2>&1 | grep "mlp dbg" | tail -3
TT_MLP_DEBUG=1 TT_ENC_MLPFUSE=1 timeout 200 python tools/dec_bench.py 9600 2>&1 | grep "mlp dbg" | tail -2
