#!/bin/bash
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
GLC_ATTN=persist timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_persist -s 2 -c 1 -f -o gpurun_out/e_attn python scripts/bench_attn.py 64 512 12 3 > gpurun_out/e_ncu.log 2>&1
tail -5 gpurun_out/e_ncu.log
