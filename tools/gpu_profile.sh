#!/bin/bash
# ncu evidence for profiles/: launch list of a short bench run + --set full captures of the dominant kernel
# (fc1+GELU, qkv, fc2+residual shapes through tools/gemm_probe.py).  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
TAG=${1:-cur}
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv \
  --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --pages-per-gpu 8 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1
echo "launch list rc=$?"
for shape in "384 1536 2 0 fc1" "384 1152 0 0 qkv" "1536 384 0 1 fc2"; do
  set -- $shape
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 3 -c 1 -f \
    -o gpurun_out/prof_${5}_$TAG python tools/gemm_probe.py $1 $2 $3 $4 0 0 5 > gpurun_out/ncu_$5_$TAG.log 2>&1
  echo "full capture $5 rc=$?"
done
ls -la gpurun_out
