#!/bin/bash
for c in 0 148 296 592 1184; do echo "== TT_ENC_CHUNK=$c"; TT_ENC_CHUNK=$c timeout 300 python tools/dec_bench.py 9600 2>&1 | grep "fused=1"; done
