#!/bin/bash
# session 3, call U: slots / group size at the head (same box)
mkdir -p gpurun_out
run() { # name, env, args
  env $2 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-configs $3 > gpurun_out/bench_u.json 2> gpurun_out/bench_u.err
  python - <<EOF
import json
d=json.load(open('gpurun_out/bench_u.json'))
print('$1:', round(d['value'],1), 'pages/s; e2e', round(d['e2e']['value'],1), 'clock', d['clocks']['sm_mhz'])
EOF
}
run "slots 3, groups of 32 pages" "TT_SLOTS=3" ""
run "slots 4, groups of 32 pages" "TT_SLOTS=4" ""
run "slots 3, groups of 64 pages" "TT_SLOTS=3" "--batch-pages 64"
