#!/bin/bash
# ncu evidence for profiles/: launch list of a short bench run + --set full captures of the dominant kernels.
# Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
TAG=${1:-cur}
export TT_BENCH_CHILD=1   # profile the bench process itself, not the supervisor's child
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 5000 --csv \
  --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --pages-per-gpu 8 --batch-pages 8 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
echo "launch list rc=$?"
for shape in "384 1536 2 0 fc1" "384 1152 0 0 qkv" "1536 384 0 1 fc2" "384 384 0 1 proj"; do
  set -- $shape
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -f \
    -o gpurun_out/prof_${5}_$TAG python tools/gemm_probe.py $1 $2 $3 $4 0 0 5 > gpurun_out/ncu_$5_$TAG.log 2>&1
  echo "full capture $5 rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 1 -c 1 -f \
    -o gpurun_out/prof_c1_2_$TAG python tools/conv_probe.py 8 1024 1024 64 64 2 > gpurun_out/ncu_c1_2_$TAG.log 2>&1
echo "full capture c1_2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 1 -c 1 -f \
    -o gpurun_out/prof_c4_$TAG python tools/conv_probe.py 8 128 128 512 512 2 > gpurun_out/ncu_c4_$TAG.log 2>&1
echo "full capture c4 rc=$?"
ls -la gpurun_out | grep $TAG
