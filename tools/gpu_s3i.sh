#!/bin/bash
# session 3, call I: how much of the decoder stage is the refinement pass
for v in 0 1 0 1; do echo "== TT_DEC_SKIP_REFINE=$v"; TT_DEC_SKIP_REFINE=$v timeout 300 python tools/dec_bench.py 2400 9600 2>&1 | grep "fused=1"; done
