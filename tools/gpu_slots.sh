#!/bin/bash
for s in 2 3 4; do for bp in 8 4; do
echo "== slots $s batch_pages $bp"; TT_SLOTS=$s timeout 300 python bench.py --no-cpu-baseline --steps 2 --warmup 2 --batch-pages $bp 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'])"
done; done
echo "== tiny gemms M=2400"; python tools/gemm_probe3.py 2400 2>&1 | grep -v Warn
