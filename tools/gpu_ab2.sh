#!/bin/bash
# engine-level knobs A/B on the bench workload (short runs)
run() { echo "== $*"; env "$@" timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-configs $EXTRA 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'clk',d['clocks']['sm_mhz'])"; }
EXTRA="" run TT_SLOTS=2
EXTRA="" run TT_SLOTS=3
EXTRA="--batch-pages 64" run TT_SLOTS=2
EXTRA="--batch-pages 16" run TT_SLOTS=2
EXTRA="" run TT_SLOTS=2 TT_CRAFT_POOLFUSE=0
