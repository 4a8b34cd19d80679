#!/bin/bash
# fused decoder bring-up: parity tests of the PARSeq path, then stage timings
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_models_gpu.py -m gpu -x -q -s -k "parseq" > gpurun_out/t_dec.log 2>&1; echo "parseq tests rc=$?"
grep -v Warning gpurun_out/t_dec.log | tail -25
timeout 300 python tools/dec_bench.py 2>&1 | grep -v Warning | tail -8
timeout 300 python tools/latency_bench.py 2>&1 | grep -v Warning | tail -9
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_dec --launch-skip 120 --launch-count 24 --csv --log-file gpurun_out/dec_launches.csv python tools/dec_bench.py 9600 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/dec_launches.csv')) if len(r)>5 and r[0].isdigit()]
for r in rows[:24]:
    print(r[4][:60], r[-1], r[-2])
PY
