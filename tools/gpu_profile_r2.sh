#!/bin/bash
# Round-2 ncu evidence for profiles/: launch list of a short bench run + --set full captures of the dominant kernels
# (LayerNorm-fused GEMM pair, fused decoder kernels, cross attention, conv1_2).  Numbers under ncu are never bench values.
mkdir -p gpurun_out
TAG=${1:-r2}
export TT_BENCH_CHILD=1   # profile the bench process itself, not the supervisor's child
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv \
  --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --pages-per-gpu 8 --batch-pages 8 --no-cpu-baseline --no-configs > gpurun_out/ncu_bench_$TAG.log 2>&1
echo "launch list rc=$?"
# producer proj-like (K 384) + consumer fc1 (GELU); producer fc2-like (K 1536) + consumer qkv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 4 -c 2 -f \
  -o gpurun_out/prof_lnpair_proj_fc1_$TAG python tools/ln_probe.py 384 1536 2 > gpurun_out/ncu_lnpair1_$TAG.log 2>&1; echo "ln pair proj+fc1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 4 -c 2 -f \
  -o gpurun_out/prof_lnpair_fc2_qkv_$TAG python tools/ln_probe.py 1536 1152 0 > gpurun_out/ncu_lnpair2_$TAG.log 2>&1; echo "ln pair fc2+qkv rc=$?"
# decoder kernels at full occupancy: the full 26-step schedule (with the early exit the launches past step ~4 are nearly empty)
for k in k_dec_dense k_dec_cross_attn; do
  TT_DEC_EARLY_EXIT=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 60 -c 2 -f \
    -o gpurun_out/prof_${k}_$TAG python tools/dec_bench.py 9600 > gpurun_out/ncu_${k}_$TAG.log 2>&1; echo "$k rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_enc_mlp -s 14 -c 1 -f \
    -o gpurun_out/prof_k_enc_mlp_$TAG python tools/dec_bench.py 9600 > gpurun_out/ncu_k_enc_mlp_$TAG.log 2>&1; echo "k_enc_mlp rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_attn_enc -s 14 -c 1 -f \
    -o gpurun_out/prof_k_attn_enc_$TAG python tools/dec_bench.py 9600 > gpurun_out/ncu_k_attn_enc_$TAG.log 2>&1; echo "k_attn_enc rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_conv1_1 -s 2 -c 1 -f \
    -o gpurun_out/prof_conv1_1_$TAG python tools/stage_bench.py 8 quick > gpurun_out/ncu_conv1_1_$TAG.log 2>&1
echo "full capture conv1_1 rc=$?"
ls -la gpurun_out | grep $TAG
