#!/bin/bash
# session 3, call R: whole-pipeline A/B of folding proj into the fused MLP kernel (same box)
mkdir -p gpurun_out
for v in 0 1 0 1; do
  TT_ENC_PROJFUSE=$v timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-configs > gpurun_out/bench_pj$v.json 2> gpurun_out/bench_pj.err
  python - <<EOF
import json
d=json.load(open('gpurun_out/bench_pj$v.json'))
print('TT_ENC_PROJFUSE=$v:', round(d['value'],1), 'pages/s; e2e', round(d['e2e']['value'],1), 'encoder ms', round(d['stages']['parseq_encoder']['ms_per_step'],1), 'clock', d['clocks']['sm_mhz'])
EOF
done
