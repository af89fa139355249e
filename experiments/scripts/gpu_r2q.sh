#!/bin/bash
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
for p in 1 0 1 0; do GLC_ATTN_POLY=$p GLC_ATTN=persist python scripts/bench_attn.py 64 512 12 50 2>&1 | tail -2 | sed "s/^/poly=$p /"; done > gpurun_out/q_attn.txt
cat gpurun_out/q_attn.txt
for p in 1 0; do GLC_ATTN_POLY=$p timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/q_bench_poly$p.json 2> gpurun_out/q_bench_poly$p.err; python -c "
import json;d=json.loads(open('gpurun_out/q_bench_poly$p.json').read().strip().splitlines()[-1]);print('poly=$p',d['value'],d['ms_per_step'],d['roofline_attention']['us_per_launch'])"; done
