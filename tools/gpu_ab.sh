#!/bin/bash
# same-box A/B of the kernel-path switches (development aid)
run() { env "$@" TT_BENCH_CHILD=1 timeout 200 python bench.py --no-cpu-baseline --steps 2 --warmup 2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['achieved']), {k:round(v['ms_per_step'],1) for k,v in d['stages'].items() if k in ('craft','parseq_encoder','parseq_decoder')})"; echo "   ^ $*"; }
run A=default
run TT_GEMM_PAIR=0
run TT_CONV_HALO=0
run TT_GEMM_TS=0
run TT_GEMM_TE=0
run TT_GEMM_EW=8
run A=default2
