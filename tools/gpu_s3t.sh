#!/bin/bash
# session 3, call T: ncu evidence of the head: launch list + --set full of the encoder block kernel and the qkv consumer
mkdir -p gpurun_out
TAG=r2g
export TT_BENCH_CHILD=1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 6000 --csv \
  --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --pages-per-gpu 8 --batch-pages 8 --no-cpu-baseline --no-configs > gpurun_out/ncu_bench_$TAG.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_enc_mlp -s 14 -c 1 -f \
    -o gpurun_out/prof_k_enc_mlp_$TAG python tools/dec_bench.py 9600 > gpurun_out/ncu_k_enc_mlp_$TAG.log 2>&1; echo "k_enc_mlp rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 4 -c 2 -f \
  -o gpurun_out/prof_lnpair_fc2_qkv_$TAG python tools/ln_probe.py 1536 1152 0 > gpurun_out/ncu_lnpair2_$TAG.log 2>&1; echo "ln pair fc2+qkv rc=$?"
ls -la gpurun_out | grep $TAG
