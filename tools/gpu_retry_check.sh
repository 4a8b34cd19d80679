#!/bin/bash
# the bench supervisor's conservative retry configuration must itself run (all round-2 fusions off, one slot)
export TT_CONV_HALO=0 TT_GEMM_TS=0 TT_GEMM_TE=0 TT_GEMM_EW=8 TT_SLOTS=1 TT_ENC_LNFUSE=0 TT_DEC_FUSED=0 TT_CRAFT_POOLFUSE=0
timeout 600 python bench.py --steps 1 --warmup 1 --total-pages 64 --no-cpu-baseline --no-configs 2>/dev/null | cut -c1-200
timeout 600 python -m pytest tests/test_e2e_gpu.py tests/test_models_gpu.py -m gpu -x -q -k "synth_page or teacher_forced or score_maps" 2>&1 | tail -2
