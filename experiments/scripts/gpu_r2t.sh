#!/bin/bash
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
: > gpurun_out/t_probe.txt
for cfg in "base 512 10 64 16" "base 1024 100 40 8"; do
  python scripts/precision_probe.py $cfg 2>&1 | grep probe >> gpurun_out/t_probe.txt
  GLC_ATTN_POLY=1 python scripts/precision_probe.py $cfg 2>&1 | grep probe >> gpurun_out/t_probe.txt
  PROBE_PRELN_F32=1 python scripts/precision_probe.py $cfg 2>&1 | grep probe >> gpurun_out/t_probe.txt
  GLC_ATTN=shift python scripts/precision_probe.py $cfg 2>&1 | grep probe >> gpurun_out/t_probe.txt
  GLC_ATTN=rows python scripts/precision_probe.py $cfg 2>&1 | grep probe >> gpurun_out/t_probe.txt
done
cat gpurun_out/t_probe.txt
