#!/bin/bash
# session 3, call J: the per-rank workload of the 8-GPU strong-scaling run (64 pages per step) on one GPU, group sizes
mkdir -p gpurun_out
for bp in 32 24 16; do
  timeout 600 python bench.py --pages-per-gpu 64 --batch-pages $bp --no-cpu-baseline --no-configs --steps 5 --warmup 3 > gpurun_out/bench_p64_b$bp.json 2> gpurun_out/bench_p64.err
  python - <<EOF
import json
d=json.load(open('gpurun_out/bench_p64_b$bp.json'))
print('batch-pages $bp:', round(d['value'],1), 'pages/s; e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],1))
EOF
done
