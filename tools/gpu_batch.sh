#!/bin/bash
timeout 300 python -m pytest tests/test_models_gpu.py tests/test_e2e_gpu.py -x -q 2>&1 | grep -E "Error|passed|failed" | cut -c1-400
for bp in 8 16 32; do
echo "== batch_pages $bp"; timeout 200 python bench.py --no-cpu-baseline --steps 2 --warmup 2 --batch-pages $bp 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], {k:round(v['ms_per_step'],1) for k,v in d['stages'].items()})"
done
