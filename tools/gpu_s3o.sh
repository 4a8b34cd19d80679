#!/bin/bash
# session 3, call O: whole-pipeline A/B of the fused encoder MLP (same box)
mkdir -p gpurun_out
for v in 0 1 0 1; do
  TT_ENC_MLPFUSE=$v timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-configs > gpurun_out/bench_mlp$v.json 2> gpurun_out/bench_mlp.err
  python - <<EOF
import json
d=json.load(open('gpurun_out/bench_mlp$v.json'))
print('TT_ENC_MLPFUSE=$v:', round(d['value'],1), 'pages/s; e2e', round(d['e2e']['value'],1), 'encoder ms', round(d['stages']['parseq_encoder']['ms_per_step'],1), 'clock', d['clocks']['sm_mhz'])
EOF
done
