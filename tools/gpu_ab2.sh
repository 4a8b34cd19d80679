#!/bin/bash
# engine-level knobs A/B on the bench workload (short runs, same box)
run() { echo "== $*"; env "$@" timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'clk',d['clocks']['sm_mhz'])"; }
run TT_SLOTS=2
run TT_SLOTS=3
run TT_SLOTS=4
run TT_SLOTS=2
run TT_SLOTS=3
