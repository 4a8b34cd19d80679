#!/bin/bash
# ncu --set full captures of the non-GEMM kernels (HBM-bound stages) from one 8-page pipeline run
mkdir -p gpurun_out
TAG=${1:-cur}
for k in k_attn_enc k_layernorm k_dec_cross_attn k_dec_attn_refine k_merge k_label_init k_crop k_page_resize k_maxpool2 k_upsample2; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f \
    -o gpurun_out/prof_${k}_$TAG python tools/stage_bench.py 8 quick > gpurun_out/ncu_${k}_$TAG.log 2>&1
  echo "$k rc=$?"
done
