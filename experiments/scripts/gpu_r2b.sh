#!/bin/bash
mkdir -p gpurun_out
export GLC_MODEL_CACHE=/tmp/glc_models
GLC_ATTN=rows GLC_ATTN_TRACE=gpurun_out/b_trace_s512.txt timeout 300 python scripts/bench_attn.py 64 512 12 1 > gpurun_out/b_trace.log 2>&1
GLC_ATTN=rows GLC_ATTN_TRACE=gpurun_out/b_trace_s1024.txt timeout 300 python scripts/bench_attn.py 16 1024 12 1 >> gpurun_out/b_trace.log 2>&1
( timeout 600 python scripts/precision_probe.py base 512 10 64 16
  GLC_ATTN=shift timeout 600 python scripts/precision_probe.py base 512 10 64 16
  GLC_ATTN=shift GLC_ATTN_C16=0 GLC_ATTN_G16=0 timeout 600 python scripts/precision_probe.py base 512 10 64 16
  PROBE_PRELN_F32=1 timeout 600 python scripts/precision_probe.py base 512 10 64 16
  timeout 900 python scripts/precision_probe.py large 1024 50 16 4
  GLC_ATTN=shift GLC_ATTN_C16=0 GLC_ATTN_G16=0 timeout 900 python scripts/precision_probe.py large 1024 50 16 4
  PROBE_PRELN_F32=1 timeout 900 python scripts/precision_probe.py large 1024 50 16 4 ) > gpurun_out/b_probe.log 2>&1
cat gpurun_out/b_trace.log gpurun_out/b_probe.log; cat gpurun_out/b_trace_s512.txt
